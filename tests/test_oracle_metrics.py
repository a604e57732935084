"""CPU: the numpy restatement of DSSIM_MP_RGB.m (oracle/metrics_oracle.py; MATLAB itself cannot run here, parity unpinned
against MATLAB) checked against the metric's own invariants and an independent direct (non-separable) evaluation of the
11 x 11 x 11 replicate-padded volume window at sampled positions."""
import numpy as np

from oracle import metrics_oracle as M


def _pair(seed, H=40, W=36):
    rs = np.random.RandomState(seed)
    a = rs.randint(0, 256, (H, W, 3)).astype(np.uint8)
    b = np.clip(a.astype(np.int32) + rs.randint(-30, 31, a.shape), 0, 255).astype(np.uint8)
    m = (rs.uniform(size=(H, W)) < 0.6).astype(np.uint8) * 255
    return a, b, m


def test_invariants():
    a, b, m = _pair(0)
    assert abs(M.dssim_mp_rgb(a, a, m)) <= 1e-15                                  # identical images: DSSIM 0
    assert abs(M.dssim_mp_rgb(a, b, m) - M.dssim_mp_rgb(b, a, m)) <= 1e-15        # symmetric
    assert 0.0 < M.dssim_mp_rgb(a, b, m) < 1.0
    assert M.matlab_ssim_map(a / 255.0, b / 255.0).shape == a.shape               # 'replicate' padding: a full-size map
    # a constant pair: variances vanish, the map is the luminance term everywhere, with either window
    ca, cb = np.full((20, 20, 3), 50, np.uint8), np.full((20, 20, 3), 200, np.uint8)
    x, y = 50 / 255.0, 200 / 255.0
    want = (1.0 - (2 * x * y + 1e-4) / (x * x + y * y + 1e-4)) / 2.0
    for w3 in (True, False):
        assert abs(M.dssim_mp_rgb(ca, cb, np.full((20, 20), 255, np.uint8), w3) - want) <= 1e-12


def test_volume_window_against_a_direct_evaluation():
    a, b, _ = _pair(1, 24, 20)
    A, B = a / 255.0, b / 255.0
    g = M._gauss()
    w3 = g[:, None, None] * g[None, :, None] * g[None, None, :]

    def direct(v, y, x, c):
        H, W, C = v.shape
        ys = np.clip(np.arange(y - 5, y + 6), 0, H - 1)
        xs = np.clip(np.arange(x - 5, x + 6), 0, W - 1)
        cs = np.clip(np.arange(c - 5, c + 6), 0, C - 1)
        return float((w3 * v[np.ix_(ys, xs, cs)]).sum())

    smap = M.matlab_ssim_map(A, B)
    for (y, x, c) in ((0, 0, 0), (3, 19, 1), (23, 7, 2), (12, 10, 1)):
        mx, my = direct(A, y, x, c), direct(B, y, x, c)
        sx, sy, sxy = direct(A * A, y, x, c) - mx * mx, direct(B * B, y, x, c) - my * my, direct(A * B, y, x, c) - mx * my
        want = ((2 * mx * my + 1e-4) * (2 * sxy + 9e-4)) / ((mx * mx + my * my + 1e-4) * (sx + sy + 9e-4))
        assert abs(smap[y, x, c] - want) <= 1e-12
