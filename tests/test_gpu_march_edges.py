"""GPU: edge cases and size-independent properties of the ray-march (C ABI: gfr_shadow_march_fwd / gfr_march_shade_fwd).
The reference has no tests of its own; these are the degenerate inputs its code can meet (empty / full face masks, the light
exactly on the image rectangle, a single sample, a non-uniform sample table, the smallest and the largest supported
image) checked against the CPU oracle, and properties that hold at BASELINE.json's full size without an oracle run."""
import numpy as np
import pytest
import torch

from oracle import relight_oracle as O

pytestmark = pytest.mark.gpu
SHADOW_TOL = 5e-6


@pytest.fixture(scope="module")
def ops():
    from geomconsistentfr_b200 import ops
    return ops


def _case(seed, H, W, B=1, noise=2.0):
    g = torch.Generator().manual_seed(seed)
    depth = torch.zeros(B, 1, H, W)
    masks = torch.zeros(B, H, W, dtype=torch.uint8)
    for b in range(B):
        d, m = O.synthetic_face(seed=seed + b, H=H, W=W, noise=noise)
        depth[b, 0] = d + 2.0 * torch.rand(H, W, generator=g)
        masks[b] = m
    return depth, masks


def _lp(L):
    return O.light_point(torch.as_tensor(L, dtype=torch.float32).view(-1, 3))[1]


def _check(ops, depth, masks, P_L, **kw):
    ref = O.shadow_march(depth, masks, P_L, inside_bonus=kw.get("inside_bonus", 0.0))
    for variant in (0, 1):
        d, _, s = ops.shadow_march_fwd(depth.cuda(), ops.mask_pack(masks.cuda()), P_L.cuda(), want_shadow=True, variant=variant, **kw)
        assert torch.equal(ref < 1e5, d.cpu() < 1e5)
        assert (s.cpu() - O.shadow_weight(ref)).abs().max() <= SHADOW_TOL
    return ref


def test_empty_mask_marks_every_pixel_unshadowed(ops):
    """mask == 0 everywhere: every sample is rejected (TRAIN:510-512) -> d_min = 1e6, shadow weight exactly 1, argmin 255."""
    depth, masks = _case(1, 64, 64)
    masks.zero_()
    for variant in (0, 1):
        d, a, s = ops.shadow_march_fwd(depth.cuda(), ops.mask_pack(masks.cuda()), _lp([(0.7, 0.1, 0.7)]).cuda(),
                                       want_argmin=True, want_shadow=True, variant=variant)
        assert bool((d == 1e6).all()) and bool((s == 1.0).all()) and bool((a == 255).all())
    _check(ops, depth, masks, _lp([(0.7, 0.1, 0.7)]))


def test_full_mask_and_smallest_and_largest_images(ops):
    for H, W in ((8, 32), (512, 512)):                        # the tile size, and the ABI's maximum (H, W <= 512)
        depth, masks = _case(H, H, W)
        masks.fill_(255)
        _check(ops, depth, masks, _lp([(-0.6, 0.4, 0.69)]), inside_bonus=5.0)
    lib_err = pytest.raises(RuntimeError)
    with lib_err:                                             # 1024 wide: refused, not silently wrong
        depth, masks = torch.zeros(1, 1, 8, 1024), torch.ones(1, 8, 1024, dtype=torch.uint8)
        ops.shadow_march_fwd(depth.cuda(), ops.mask_pack(masks.cuda()), _lp([(0.0, 0.0, 1.0)]).cuda())
    with lib_err:                                             # width not a multiple of the 32-pixel tile
        depth, masks = torch.zeros(1, 1, 64, 48), torch.ones(1, 64, 48, dtype=torch.uint8)
        ops.shadow_march_fwd(depth.cuda(), ops.mask_pack(masks.cuda()), _lp([(0.0, 0.0, 1.0)]).cuda())


@pytest.mark.parametrize("Lxy", [(127.0, 0.0), (-128.0, 5.0), (3.0, 128.0), (-7.0, -127.0), (127.0, 128.0), (127.5, 0.0),
                                 (-128.5, -127.5)])
def test_light_exactly_on_and_just_off_the_image_rectangle(ops, Lxy):
    """The 9-way end-point branch (TRAIN:383-442) and the +5 bonus (TEST1:495) switch on `<` / `<=` against -128, 127,
    -127, 128: light points placed exactly on those values and half a pixel outside."""
    depth, masks = _case(7, 256, 256)
    P_L = torch.tensor([[Lxy[0], Lxy[1], 3000.0]])
    _check(ops, depth, masks, P_L, inside_bonus=5.0)


def test_single_sample_and_non_uniform_table(ops):
    depth, masks = _case(3, 64, 96)
    P_L = _lp([(0.5, -0.5, 0.70710678)])
    bits = ops.mask_pack(masks.cuda())
    xx, yy = O.pixel_grid(64, 96)
    for t in (np.array([0.31]), np.array([0.02, 0.05, 0.051, 0.3, 0.9, 0.97])):      # non-uniform: culling is bypassed
        ref, _ = O._march_one(depth[0, 0], masks[0], P_L[0], xx, yy, torch.from_numpy(t), 0.0, 32)
        for variant in (0, 1):
            d, a, _ = ops.shadow_march_fwd(depth.cuda(), bits, P_L.cuda(), samples=t, want_argmin=True, variant=variant)
            fin = ref < 1e5
            assert torch.equal(fin, d[0].cpu() < 1e5)
            assert (d[0].cpu() - ref)[fin].abs().max() <= 2e-5 * max(1.0, float(ref[fin].max()))
            assert int(a.max()) in (255, *range(len(t))) and int(a[a < 255].max()) < len(t)


def test_lighting_transfer_table_and_bonus_rectangle(ops):
    """159 samples from 0.03 (TEST_LT:451) and the +-4-image-size rule (TEST_LT:503) through the operator."""
    depth, masks = _case(11, 256, 256)
    t = np.arange(0.03, 0.825, 0.005)
    rect = (-1024.0, 1024.0, 4.0 * (1 - 256), 1024.0)
    for L, bonus_applies in (((0.12, -0.2, 0.9724), True), ((-0.7574, 0.0, 0.6529), False)):
        P_L = _lp([L])
        ref = O.shadow_march(depth, masks, P_L, t0=O.LT["t0"], n=O.LT["num_samples"], inside_bonus=5.0, inside_rule="wide")
        plain = O.shadow_march(depth, masks, P_L, t0=O.LT["t0"], n=O.LT["num_samples"], inside_bonus=0.0)
        assert float(((ref - plain)[plain < 1e5] - (5.0 if bonus_applies else 0.0)).abs().max()) <= 1e-5
        for variant in (0, 1):
            d, _, _ = ops.shadow_march_fwd(depth.cuda(), ops.mask_pack(masks.cuda()), P_L.cuda(), samples=t, inside_bonus=5.0,
                                           variant=variant, bonus_rect=rect)
            fin = ref < 1e5
            assert (d.cpu() - ref)[fin].abs().max() <= 2e-5 * max(1.0, float(ref[fin].max()))


def test_full_size_properties_without_an_oracle(ops):
    """BASELINE configs[1] size (B = 8, 256x256), size-independent properties:
    (1) 0 <= shadow <= 1 and shadow == 1 exactly where no sample hit the face;
    (2) growing the face mask can only lower d_min (the min runs over a superset of samples);
    (3) a batch equals its images marched one by one (no cross-image state), and L lights per face equals L batches;
    (4) both kernel variants and the fused march+shade launch give the same bits."""
    from geomconsistentfr_b200.synthetic import LIGHTS_18
    B = 8
    depth, masks = _case(21, 256, 256, B=B)
    P_L = _lp(LIGHTS_18[4:4 + B])
    bits = ops.mask_pack(masks.cuda())
    d, a, s = ops.shadow_march_fwd(depth.cuda(), bits, P_L.cuda(), want_argmin=True, want_shadow=True)
    # (s = 1 - 4e/(1+e)^2 in fp32 can come out one ulp below 0 where d_min ~ 1e-6: the reference's formula does the same)
    assert float(s.min()) >= -2e-7, float(s.min())
    assert float(s.max()) <= 1.0, float(s.max())
    assert bool((s[a == 255] == 1.0).all()) and bool((d[a == 255] == 1e6).all())
    grown = torch.nn.functional.max_pool2d(masks.float()[:, None], 9, 1, 4)[:, 0].to(torch.uint8)
    d2, _, _ = ops.shadow_march_fwd(depth.cuda(), ops.mask_pack(grown.cuda()), P_L.cuda())
    assert bool((d2 <= d).all()), int((d2 > d).sum())
    assert bool((d2 < d).any())
    for b in (0, 5):
        db, _, _ = ops.shadow_march_fwd(depth[b:b + 1].cuda(), ops.mask_pack(masks[b:b + 1].cuda()), P_L[b:b + 1].cuda())
        assert torch.equal(db[0], d[b]), b
    dl, _, _ = ops.shadow_march_fwd(depth[:2].cuda(), ops.mask_pack(masks[:2].cuda()), P_L.cuda())      # 2 faces x 4 lights
    for f in range(2):
        for l in range(4):
            one, _, _ = ops.shadow_march_fwd(depth[f:f + 1].cuda(), ops.mask_pack(masks[f:f + 1].cuda()), P_L[4 * f + l:4 * f + l + 1].cuda())
            assert torch.equal(one[0], dl[4 * f + l]), (f, l)
    d1, _, _ = ops.shadow_march_fwd(depth.cuda(), bits, P_L.cuda(), variant=1)
    assert torch.equal(d1, d), "variant 1 != variant 0"
    amb = torch.full((B,), 0.3, device="cuda")
    o = ops.march_shade_fwd(torch.rand(B, 3, 256, 256, device="cuda"), depth.cuda(), bits, P_L.cuda(), amb,
                            want=("shadow", "d_min"))
    assert torch.equal(o["d_min"], d), "fused d_min"
    assert torch.equal(o["shadow"], s), "fused shadow"
