"""GPU: the output stage of the inference drivers (csrc/postprocess.cu, geomconsistentfr_b200/inference.py) against the
numpy oracle (oracle/postprocess_oracle.py) — bit-exact, it is integer/byte output — and, end to end, against the 10
relit PNGs the reference ships, on EVERY pixel."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def ffhq():
    return np.load(os.path.join(G, "ffhq.npz"))


@pytest.fixture(scope="module")
def net():
    from geomconsistentfr_b200 import RelightNet
    n = RelightNet()
    n.load_state_dict(torch.load(os.path.join(G, "model_epoch99.pth"), map_location="cpu"), strict=True)
    return n.cuda().eval()


def _rand_case(seed, B, H, W, per_face_mask):
    g = np.random.default_rng(seed)
    image = g.random((B, H, W, 3))                                       # f64 like the reference
    image[0, :4] = np.round(image[0, :4] * 255) / 255 + 0.5 / 255        # exact .5 ties after * 255
    rendered = (g.random((B, 3, H, W)) * 1.3 - 0.1).astype(np.float32)   # out of range on both sides: saturation
    rendered[0, :, 5, :8] = (np.arange(8, dtype=np.float32) + 0.5) / 255.0
    mask = g.choice(np.array([0, 64, 128, 255], np.uint8), size=((B if per_face_mask else 1), H, W), p=[0.4, 0.1, 0.1, 0.4])
    return image, rendered, mask


@pytest.mark.parametrize("B,H,W,per_face", [(1, 256, 256, False), (3, 64, 96, True), (2, 40, 72, False)])
@pytest.mark.parametrize("f64", [True, False])
def test_composite_is_bit_exact(B, H, W, per_face, f64):
    from geomconsistentfr_b200 import ops
    from oracle import postprocess_oracle as P
    image, rendered, mask = _rand_case(B * 7 + H, B, H, W, per_face)
    if not f64:
        image = image.astype(np.float32)
    got = ops.composite_bgr_u8(torch.from_numpy(image).cuda(), torch.from_numpy(rendered).cuda(),
                               torch.from_numpy(mask if per_face else mask[0]).cuda()).cpu().numpy()
    for b in range(B):
        want = P.composite_bgr_u8(image[b], rendered[b], mask[b if per_face else 0])
        assert np.array_equal(got[b], want), b


@pytest.mark.parametrize("H,W", [(256, 256), (48, 80)])
def test_export_planes_is_bit_exact(H, W):
    from geomconsistentfr_b200 import ops
    from oracle import postprocess_oracle as P
    g = np.random.default_rng(H)
    albedo = g.random((1, 3, H, W)).astype(np.float32)
    depth = (g.standard_normal((1, 1, H, W)) * 60 + 40).astype(np.float32)
    shadow = g.random((1, H, W)).astype(np.float32)
    final = (g.random((1, H, W)) * 1.2).astype(np.float32)
    normals = (g.random((1, 3, H, W)) * 2 - 1).astype(np.float32)
    mask = g.choice(np.array([0, 64, 128, 255], np.uint8), size=(H, W))
    c = lambda a: torch.from_numpy(a).cuda()
    got = ops.export_planes_u8(c(albedo), c(depth), c(shadow), c(final), c(normals), c(mask))
    want = P.export_planes_u8(albedo[0], depth[0], shadow[0], final[0], normals[0], mask)
    assert set(got) == set(want)
    for k in want:
        assert np.array_equal(got[k][0].cpu().numpy(), want[k]), k
    only = ops.export_planes_u8(None, c(depth), None, None, None, c(mask), want=("depth",))
    assert list(only) == ["depth"] and torch.equal(only["depth"], got["depth"])


@pytest.mark.parametrize("max_sum", [29, 30])
@pytest.mark.parametrize("C", [3, 1])
def test_border_median_fix_is_bit_exact(ffhq, max_sum, C):
    from geomconsistentfr_b200 import ops
    from oracle import postprocess_oracle as P
    g = np.random.default_rng(3)
    masks = ffhq["masks"][:3]
    img = g.integers(0, 256, size=(3, 256, 256, C), dtype=np.uint8)
    got = ops.border_median_fix_u8(torch.from_numpy(img).cuda(), torch.from_numpy(masks).cuda(), max_sum).cpu().numpy()
    for b in range(3):
        assert np.array_equal(got[b], P.border_fix_u8(img[b], masks[b], max_sum)), b
    # odd size, mask touching the image edge (zero padding of both the box sum and the median)
    m = np.zeros((40, 72), np.uint8); m[:9, :11] = 255; m[30:, 60:] = 128; m[20, 20] = 64
    im = g.integers(0, 256, size=(1, 40, 72, C), dtype=np.uint8)
    got = ops.border_median_fix_u8(torch.from_numpy(im).cuda(), torch.from_numpy(m).cuda(), max_sum).cpu().numpy()
    assert np.array_equal(got[0], P.border_fix_u8(im[0], m, max_sum))
    assert (got[0] != im[0]).any()


def test_shipped_pngs_every_pixel_through_the_driver(ffhq, net):
    """relight(..., fix_border=True) = TEST1 main() + fix_border_artifacts_CVPR2022.m, batched: all 10 shipped PNGs,
    every pixel, <= 1 grey level (same bar the oracle meets in tests/test_oracle_golden.py)."""
    from geomconsistentfr_b200 import relight
    r = relight(net, ffhq["q"] / 1020.0, ffhq["masks"], ffhq["lights"], fix_border=True)["rendered_image"]
    want = ffhq["pngs_bgr"].astype(np.int32)
    d = np.abs(r.astype(np.int32) - want)
    assert d.max() <= 1, [int(d[i].max()) for i in range(10)]
    assert d.mean() < 0.03
    raw = relight(net, ffhq["q"] / 1020.0, ffhq["masks"], ffhq["lights"])["rendered_image"]
    assert (np.abs(raw.astype(np.int32) - want).max(3) > 1).sum() > 10000          # without the fix: thousands off


def test_single_image_driver_and_planes_match_oracle_postprocess(ffhq, net):
    """relight_single_image == the oracle output stage applied to the module's own forward outputs (bit-exact), and the
    batch-of-one planes follow TESTB:590-608."""
    from geomconsistentfr_b200 import intrinsic_matrix, relight, relight_single_image
    from oracle import postprocess_oracle as P
    i = 3
    img, mask, light = ffhq["q"][i] / 1020.0, ffhq["masks"][i], ffhq["lights"][i]
    res = relight(net, img, mask, light, planes=True)
    one = relight_single_image(net, img, mask, light)
    assert np.array_equal(one, res["rendered_image"][0])
    with torch.no_grad():
        o = net(torch.from_numpy(img).float()[None].cuda(), 200, intrinsic_matrix().cuda(),
                (torch.from_numpy(mask).double() / 255.0).view(256, 256, 1).cuda(), torch.from_numpy(light).view(1, 3, 1, 1).cuda(),
                torch.full((1, 1, 1), 0.5).cuda(), None)
    n = lambda t: t[0].cpu().numpy()
    assert np.array_equal(one, P.composite_bgr_u8(img, n(o[5]), mask))
    want = P.export_planes_u8(n(o[0]), n(o[1]), n(o[2]), n(o[8]), n(o[9]), mask)
    for k, v in want.items():
        assert np.array_equal(res[k][0], v), k


def test_driver_rejects_cpu_and_train_mode(net):
    from geomconsistentfr_b200 import RelightNet, relight
    with pytest.raises(RuntimeError):
        relight(RelightNet().eval(), np.zeros((256, 256, 3)), np.zeros((256, 256), np.uint8), (0, 0, 1))
    with pytest.raises(RuntimeError):
        relight(net, np.zeros((256, 256, 3)), np.zeros((256, 256), np.float64), (0, 0, 1))   # /255 float masks are refused


def test_masked_mse_metric_matches_the_matlab_expression(ffhq):
    """MSE_MP.m:15-25 on the device vs its numpy restatement (fp64 sums; tolerance = summation order only)."""
    from geomconsistentfr_b200 import ops
    from oracle import postprocess_oracle as P
    g = np.random.default_rng(5)
    recon = ffhq["pngs_bgr"][:4]
    gt = np.clip(recon.astype(np.int32) + g.integers(-20, 21, recon.shape), 0, 255).astype(np.uint8)
    masks = ffhq["masks"][:4]
    c = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    got = ops.masked_mse_u8(c(recon), c(gt), c(masks)).cpu().numpy()
    for b in range(4):
        want = P.masked_mse(recon[b], gt[b], masks[b])
        assert abs(got[b] - want) <= 1e-10 * want, b
    shared = ops.masked_mse_u8(c(recon), c(gt), c(masks[0])).cpu().numpy()
    assert abs(shared[2] - P.masked_mse(recon[2], gt[2], masks[0])) <= 1e-10 * shared[2]
    assert float(ops.masked_mse_u8(c(recon), c(recon), c(masks))[0]) == 0.0


@pytest.mark.parametrize("window_3d", [True, False])
def test_masked_dssim_metric_vs_numpy_restatement(window_3d):
    """gfr_masked_ssim_u8 (DSSIM_MP_RGB.m:15-27, MATLAB's volume-window ssim map, fp64) against oracle/metrics_oracle.py on
    random 8-bit image pairs incl. a ragged size and per-image / shared masks."""
    from geomconsistentfr_b200 import ops
    from oracle import metrics_oracle as M
    rs = np.random.RandomState(5)
    for (B, H, W, shared) in ((3, 256, 256, True), (2, 70, 45, False)):
        a = rs.randint(0, 256, (B, H, W, 3)).astype(np.uint8)
        b = np.clip(a.astype(np.int32) + rs.randint(-40, 41, a.shape), 0, 255).astype(np.uint8)
        m = ((rs.uniform(size=(1 if shared else B, H, W)) < 0.6) * int(rs.choice([64, 128, 255]))).astype(np.uint8)
        got = ops.masked_dssim_u8(torch.from_numpy(a).cuda(), torch.from_numpy(b).cuda(), torch.from_numpy(m[0] if shared else m).cuda(),
                                  window_3d=window_3d).cpu().numpy()
        want = np.array([M.dssim_mp_rgb(a[i], b[i], m[0] if shared else m[i], window_3d) for i in range(B)])
        assert np.abs(got - want).max() <= 1e-10, (got, want)
    same = ops.masked_dssim_u8(torch.from_numpy(a).cuda(), torch.from_numpy(a).cuda(), torch.from_numpy(m).cuda(), window_3d=window_3d)
    assert float(same.abs().max()) <= 1e-14
