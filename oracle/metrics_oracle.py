"""TEST INFRASTRUCTURE — not product code.

numpy / scipy restatement of the reference's MATLAB evaluation metric DSSIM_MP_RGB.m:15-27.  No MATLAB or Octave exists in
this image, so MATLAB's `ssim` itself cannot be executed: PARITY UNPINNED against MATLAB; what is restated is MATLAB's
documented algorithm for `[ssimval, ssimmap] = ssim(A, ref)` with default arguments on a double M x N x 3 array:

  * the array is filtered as a volume with an isotropic Gaussian, sigma 1.5, radius ceil(3 sigma) = 5 -> 11 x 11 x 11 taps,
    `imfilter(..., 'conv', 'replicate')` (scipy.ndimage.correlate1d(mode='nearest') per axis: the kernel is symmetric);
  * DynamicRange 1 for doubles: C1 = (0.01)^2, C2 = (0.03)^2, exponents [1 1 1] ->
    map = ((2 mu_x mu_y + C1)(2 s_xy + C2)) / ((mu_x^2 + mu_y^2 + C1)(s_x^2 + s_y^2 + C2)), s from filt(x^2) - mu^2;
  * DSSIM_MP_RGB.m:24-26: average = sum(map * mask3) / sum(mask3) with mask = uint8 mask / 255, DSSIM = (1 - average) / 2.
"""
import numpy as np
from scipy.ndimage import correlate1d


def _gauss(sigma=1.5):
    r = int(np.ceil(3 * sigma))
    c = np.arange(-r, r + 1, dtype=np.float64)
    g = np.exp(-(c ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def matlab_ssim_map(A, ref, window_3d=True):
    """A, ref: [H,W,3] float64 in [0,1] -> the full-size SSIM map [H,W,3]."""
    g = _gauss()
    axes = (0, 1, 2) if window_3d else (0, 1)

    def filt(v):
        for ax in axes:
            v = correlate1d(v, g, axis=ax, mode="nearest")
        return v

    C1, C2 = 0.01 ** 2, 0.03 ** 2
    mux, muy = filt(A), filt(ref)
    sx, sy, sxy = filt(A * A) - mux * mux, filt(ref * ref) - muy * muy, filt(A * ref) - mux * muy
    return ((2 * mux * muy + C1) * (2 * sxy + C2)) / ((mux * mux + muy * muy + C1) * (sx + sy + C2))


def dssim_mp_rgb(recon_u8, gt_u8, mask_u8, window_3d=True):
    """DSSIM_MP_RGB.m:15-27 for one image: uint8 [H,W,3] x 2, uint8 mask [H,W] -> DSSIM."""
    A, ref = recon_u8.astype(np.float64) / 255.0, gt_u8.astype(np.float64) / 255.0
    m3 = np.repeat((mask_u8.astype(np.float64) / 255.0)[:, :, None], 3, axis=2)
    smap = matlab_ssim_map(A, ref, window_3d)
    return (1.0 - (smap * m3).sum() / m3.sum()) / 2.0
