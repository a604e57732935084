"""An INDEPENDENT float64 implementation of the SSIM that `pytorch_msssim.ssim` computes (call site TRAIN:643), written
from the published formula (Wang et al. 2004 as parameterised in the pytorch_msssim README: 11-tap Gaussian window,
sigma 1.5, K = (0.01, 0.03), 'valid' filtering, mean over the map, optional ReLU, mean over planes) with
scipy.ndimage.correlate1d — no code shared with oracle/relight_oracle.py or the CUDA kernel.  Used to pin both."""
import numpy as np
from scipy.ndimage import correlate1d


def gaussian_window(size=11, sigma=1.5):
    c = np.arange(size, dtype=np.float64) - size // 2
    g = np.exp(-(c ** 2) / (2.0 * sigma ** 2))
    return g / g.sum()


def _blur_valid(a, g):
    """Separable Gaussian over the last two axes, 'valid' region only."""
    r = len(g) // 2
    a = correlate1d(a, g, axis=-2, mode="constant", cval=0.0)
    a = correlate1d(a, g, axis=-1, mode="constant", cval=0.0)
    return a[..., r:a.shape[-2] - r, r:a.shape[-1] - r]


def ssim_planes(X, Y, data_range=1.0, K=(0.01, 0.03), size=11, sigma=1.5):
    """X, Y [N,C,H,W] -> (per-plane mean SSIM [N,C] float64, the SSIM map [N,C,H-10,W-10])."""
    X, Y = np.asarray(X, np.float64), np.asarray(Y, np.float64)
    g = gaussian_window(size, sigma)
    C1, C2 = (K[0] * data_range) ** 2, (K[1] * data_range) ** 2
    mx, my = _blur_valid(X, g), _blur_valid(Y, g)
    vx = _blur_valid(X * X, g) - mx * mx
    vy = _blur_valid(Y * Y, g) - my * my
    cxy = _blur_valid(X * Y, g) - mx * my
    smap = ((2 * mx * my + C1) * (2 * cxy + C2)) / ((mx * mx + my * my + C1) * (vx + vy + C2))
    return smap.mean(axis=(-2, -1)), smap


def ssim(X, Y, data_range=1.0, nonnegative_ssim=True):
    per, _ = ssim_planes(X, Y, data_range)
    if nonnegative_ssim:
        per = np.maximum(per, 0.0)
    return float(per.mean())
