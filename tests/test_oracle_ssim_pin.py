"""CPU: pins the oracle's `ssim` (the restatement of the un-vendored pytorch_msssim.ssim, call site TRAIN:643) against an
independent float64 scipy implementation of the published formula (tests/_ssim_scipy.py) and against the metric's
published invariants and closed forms."""
import numpy as np
import torch

from _ssim_scipy import gaussian_window, ssim, ssim_planes
from oracle import relight_oracle as O


def _pair(seed, N=2, C=3, S=64):
    rs = np.random.RandomState(seed)
    x = rs.uniform(0, 1, (N, C, S, S))
    y = np.clip(x + 0.15 * rs.standard_normal(x.shape), 0, 1)
    return x, y


def test_oracle_ssim_equals_independent_scipy_implementation():
    for seed, S in ((0, 64), (1, 256), (2, 33)):
        x, y = _pair(seed, S=S)
        want = ssim(x, y, 1.0, nonnegative_ssim=True)
        got32 = float(O.ssim(torch.from_numpy(x).float(), torch.from_numpy(y).float(), data_range=1.0, size_average=True,
                             nonnegative_ssim=True))
        got64 = float(O.ssim(torch.from_numpy(x), torch.from_numpy(y), data_range=1.0, size_average=True, nonnegative_ssim=True))
        assert abs(got64 - want) <= 1e-7, (seed, got64, want)       # fp64 vs fp64 (the window itself is built in fp32 upstream)
        assert abs(got32 - want) <= 2e-6, (seed, got32, want)
        per = O.ssim(torch.from_numpy(x), torch.from_numpy(y), data_range=1.0, size_average=False, nonnegative_ssim=False)
        assert np.abs(per.numpy() - ssim_planes(x, y)[0].mean(1)).max() <= 1e-7


def test_published_invariants_and_closed_forms():
    x, y = _pair(3, S=48)
    tx, ty = torch.from_numpy(x), torch.from_numpy(y)
    kw = dict(data_range=1.0, size_average=True, nonnegative_ssim=True)
    assert abs(float(O.ssim(tx, tx, **kw)) - 1.0) <= 1e-12                                  # ssim(x, x) = 1
    assert abs(float(O.ssim(tx, ty, **kw)) - float(O.ssim(ty, tx, **kw))) <= 1e-12          # symmetry
    assert float(O.ssim(tx, ty, **kw)) < 1.0
    _, smap = ssim_planes(np.zeros((1, 1, 256, 256)), np.zeros((1, 1, 256, 256)))
    assert smap.shape[-2:] == (246, 246)                                                    # 'valid' window: 256 - 10
    # constant images: variances vanish, SSIM = (2ab + C1) / (a^2 + b^2 + C1)
    a, b = 0.3, 0.7
    want = (2 * a * b + 1e-4) / (a * a + b * b + 1e-4)
    got = float(O.ssim(torch.full((1, 3, 32, 32), a, dtype=torch.float64), torch.full((1, 3, 32, 32), b, dtype=torch.float64), **kw))
    assert abs(got - want) <= 2e-5          # the upstream window is normalised in fp32 (sum = 1 + 6e-8): variance terms of ~1e-8 against C2 = 9e-4
    # anti-correlated textures give a negative plane mean: nonnegative_ssim clamps the PLANE MEAN (not the map) at 0
    rs = np.random.RandomState(4)
    t = rs.uniform(0, 1, (1, 1, 40, 40))
    neg = O.ssim(torch.from_numpy(t), torch.from_numpy(1.0 - t), data_range=1.0, size_average=True, nonnegative_ssim=False)
    assert float(neg) < 0 and float(O.ssim(torch.from_numpy(t), torch.from_numpy(1.0 - t), **kw)) == 0.0
    g = gaussian_window()
    assert abs(g.sum() - 1.0) < 1e-15 and g.argmax() == 5 and np.allclose(g, g[::-1])
