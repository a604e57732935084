"""GPU parity: CUDA ray-march and shade/render (through the C ABI) vs the reference's own outputs
(tests/golden, generated from the unmodified reference) and vs the CPU oracle on seeded inputs."""
import os

import numpy as np
import pytest
import torch

from oracle import relight_oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")

# fp32 tolerance on the shadow weight s = tanh^2(d/2) in [0,1]: positions/bilinear are fp64 exactly as the
# reference; the fp32 tail (cross product, sqrt, exp) differs from torch by a few ulp.
SHADOW_TOL = 5e-6
TAGS = ["right_mid", "right_above", "right_below", "left_mid", "left_above", "left_below",
        "mid_above", "mid_below", "inside", "grazing"]


@pytest.fixture(scope="module")
def ops():
    from geomconsistentfr_b200 import ops
    return ops


@pytest.fixture(scope="module")
def march():
    return np.load(os.path.join(G, "march.npz"))


def _light_pt(L):
    return O.light_point(torch.as_tensor(L, dtype=torch.float32).view(-1, 3))[1]


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("tag", TAGS)
def test_march_vs_reference_golden(ops, march, tag, variant):
    depth = torch.from_numpy(march["depth"]).view(1, 1, 256, 256).cuda()
    bits = ops.mask_pack(torch.from_numpy(march["mask_u8"]).view(1, 256, 256).cuda())
    P_L = _light_pt(march["light_" + tag]).cuda()
    _, _, s = ops.shadow_march_fwd(depth, bits, P_L, inside_bonus=5.0, want_shadow=True, variant=variant)
    diff = np.abs(s[0].cpu().numpy() - march["shadow_" + tag])
    assert diff.max() <= SHADOW_TOL, (tag, diff.max(), int((diff > SHADOW_TOL).sum()))


@pytest.mark.parametrize("variant", [0, 1, 2])
@pytest.mark.parametrize("H,W,B", [(64, 64, 3), (96, 128, 2), (256, 256, 2)])
def test_march_vs_oracle_seeded(ops, H, W, B, variant):
    """Per-image masks (TRAIN:510), argmin, d_min, odd sizes; lights from all quadrants incl. inside."""
    g = torch.Generator().manual_seed(1234 + H + W)
    lights = [(0.7, 0.3, 0.6), (-0.4, -0.6, 0.7), (0.02, 0.01, 0.999)][:B]
    depth = torch.zeros(B, 1, H, W)
    masks = torch.zeros(B, H, W, dtype=torch.uint8)
    for b in range(B):
        d, m = O.synthetic_face(seed=b + H, H=H, W=W, noise=2.0)
        depth[b, 0] = d + 3.0 * torch.rand(H, W, generator=g)
        masks[b] = m
    P_L = _light_pt(lights)
    d_ref, a_ref = O.shadow_march(depth, masks, P_L, inside_bonus=5.0, return_argmin=True)
    bits = ops.mask_pack(masks.cuda())
    d, a, s = ops.shadow_march_fwd(depth.cuda(), bits, P_L.cuda(), inside_bonus=5.0, want_argmin=True,
                                   want_shadow=True, variant=variant)
    d, a, s = d.cpu(), a.cpu(), s.cpu()
    finite = d_ref < 1e5
    assert torch.equal(finite, d < 1e5)
    assert (d - d_ref)[finite].abs().max() <= 2e-5 * max(1.0, float(d_ref[finite].max()))
    assert (s - O.shadow_weight(d_ref)).abs().max() <= SHADOW_TOL
    # argmin: identical except at exact fp32 ties between neighbouring samples
    a_ref = torch.where(finite, a_ref, torch.full_like(a_ref, 255))
    assert (a.long() != a_ref).float().mean() < 1e-3


def test_march_shared_mask_equals_per_image_mask(ops, march):
    depth = torch.from_numpy(march["depth"]).view(1, 1, 256, 256).repeat(3, 1, 1, 1).cuda()
    m = torch.from_numpy(march["mask_u8"]).view(1, 256, 256).cuda()
    P_L = _light_pt([march["light_" + t] for t in TAGS[:3]]).cuda()
    d1, _, _ = ops.shadow_march_fwd(depth, ops.mask_pack(m), P_L)
    d3, _, _ = ops.shadow_march_fwd(depth, ops.mask_pack(m.repeat(3, 1, 1)), P_L)
    assert torch.equal(d1, d3)


def test_mask_pack_dtypes(ops, march):
    m = torch.from_numpy(march["mask_u8"]).view(1, 256, 256)
    ref = ops.mask_pack(m.cuda())
    for t in (m.float() / 255.0, m.double() / 255.0, m > 0):
        assert torch.equal(ops.mask_pack(t.cuda()), ref)
    flat = (m.view(-1) != 0).numpy()
    row = ref.cpu().numpy()[0]
    words = row[:256 * 256 // 32].view(np.uint32)
    bits = ((words[:, None] >> np.arange(32, dtype=np.uint32)[None]) & 1).astype(bool).reshape(-1)
    assert np.array_equal(bits, flat)
    rr, cc = np.nonzero(march["mask_u8"])                       # the 4 extra words: bounding box {c_lo, -c_hi, r_lo, -r_hi}
    assert list(row[-4:]) == [cc.min(), -cc.max(), rr.min(), -rr.max()]
    empty = ops.mask_pack(torch.zeros(1, 64, 64, dtype=torch.uint8, device="cuda")).cpu().numpy()[0]
    assert not empty[:-4].any() and empty[-4] > -empty[-3]       # empty mask: c_lo > c_hi


def test_march_culling_is_exact_on_offcentre_masks(ops):
    """Sample-range culling against the mask bounding box (variant 0) vs the cull-free literal kernel (variant 1):
    bit-identical d_min / argmin for small, off-centre, ragged and empty masks and lights on every side."""
    H = W = 128
    g = torch.Generator().manual_seed(21)
    depth = (torch.rand(4, 1, H, W, generator=g) * 40.0).cuda()
    masks = torch.zeros(4, H, W, dtype=torch.uint8)
    masks[0, 10:30, 90:120] = 1
    masks[1, 100:128, 0:25] = 255
    masks[2] = (torch.rand(H, W, generator=g) > 0.97).to(torch.uint8)
    bits = ops.mask_pack(masks.cuda())                                                       # mask 3 stays empty
    L = torch.tensor([(0.7, 0.1, 0.7), (-0.6, -0.5, 0.62), (0.004, 0.003, 1.0), (0.0, 0.7071, 0.7071)])
    P_L = (4013.0 * torch.nn.functional.normalize(L, dim=1)).cuda()
    d0, a0, _ = ops.shadow_march_fwd(depth, bits, P_L, inside_bonus=5.0, want_argmin=True, variant=0)
    d1, a1, _ = ops.shadow_march_fwd(depth, bits, P_L, inside_bonus=5.0, want_argmin=True, variant=1)
    assert torch.equal(d0, d1) and torch.equal(a0, a1)
    assert float(d0[3].min()) >= 1e6                                                         # empty mask: every ray misses
    # variant 2 = the warp-per-ray mapping of the north-star sketch (A/B only): the same bits, arg-min ties included
    d2, a2, _ = ops.shadow_march_fwd(depth, bits, P_L, inside_bonus=5.0, want_argmin=True, variant=2)
    assert torch.equal(d0, d2) and torch.equal(a0, a2)


def test_coarse_group_skip_is_exact_on_masks_with_holes(ops):
    """The 8x8-block occupancy map lets a warp skip groups of 4 samples whose nearest pixels are provably outside the face
    (variant 0); variant 1 tests every sample.  Masks with holes, thin bridges, isolated pixels and a frame around the image
    border, 256x256 like the bench: bit-identical."""
    H = W = 256
    g = torch.Generator().manual_seed(5)
    depth = (torch.rand(4, 1, H, W, generator=g) * 60.0).cuda()
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    masks = torch.zeros(4, H, W, dtype=torch.uint8)
    masks[0] = ((((xx - 128) / 80.0) ** 2 + ((yy - 128) / 100.0) ** 2) < 1.0).to(torch.uint8)
    masks[0, 90:150, 100:160] = 0                                        # a big hole
    masks[0, 118:122, 100:160] = 1                                       # a thin bridge through it
    masks[1] = (torch.rand(H, W, generator=g) > 0.9995).to(torch.uint8)  # isolated pixels
    masks[2, :3, :] = 1; masks[2, -3:, :] = 1; masks[2, :, :3] = 1; masks[2, :, -3:] = 1      # a frame on the border
    masks[3, 40:200:16, 40:200] = 1                                      # stripes one pixel thick, 16 apart
    bits = ops.mask_pack(masks.cuda())
    for L in ((0.7518, 0.0, 0.6594), (-0.5151, 0.4722, 0.7154), (0.01, 0.02, 0.9997), (0.8138, -0.3420, 0.4698)):
        P_L = (4013.0 * torch.nn.functional.normalize(torch.tensor([L] * 4), dim=1)).cuda()
        d0, a0, _ = ops.shadow_march_fwd(depth, bits, P_L, inside_bonus=5.0, want_argmin=True, variant=0)
        d1, a1, _ = ops.shadow_march_fwd(depth, bits, P_L, inside_bonus=5.0, want_argmin=True, variant=1)
        assert torch.equal(d0, d1) and torch.equal(a0, a1), L


def test_warp_shape_and_sample_pairs_are_bit_identical(ops):
    """Round 2: the default kernel's A/B configurations — 32x1 vs 8x4 pixels per warp, samples one by one vs in groups of 2 / 3 / 4 (the
    group's other samples are evaluated speculatively), tile-major vs light-aware CTA order — against the literal kernel: bit-identical d_min / arg-min / fused
    shading on ragged masks, an odd sample count (the pair loop's tail) and lights on every side."""
    H = W = 128
    g = torch.Generator().manual_seed(33)
    depth = (torch.rand(4, 1, H, W, generator=g) * 50.0).cuda()
    yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
    masks = torch.zeros(4, H, W, dtype=torch.uint8)
    masks[0] = ((((xx - 64) / 40.0) ** 2 + ((yy - 64) / 50.0) ** 2) < 1.0).to(torch.uint8)
    masks[1, 5:60, 70:128] = 1
    masks[2] = (torch.rand(H, W, generator=g) > 0.9).to(torch.uint8)
    masks[3, 64:65, :] = 1
    bits = ops.mask_pack(masks.cuda())
    albedo = torch.rand(4, 3, H, W, generator=g).cuda()
    amb = torch.full((4,), 0.3, device="cuda")
    L = torch.tensor([(0.7, 0.1, 0.7), (-0.6, -0.5, 0.62), (0.004, 0.003, 1.0), (0.0, 0.7071, 0.7071)])
    P_L = (4013.0 * torch.nn.functional.normalize(L, dim=1)).cuda()
    try:
        for t in (None, np.arange(0.025, 0.825, 0.005)[:157]):
            d1, a1, _ = ops.shadow_march_fwd(depth, bits, P_L, samples=t, inside_bonus=5.0, want_argmin=True, variant=1)
            ref = None
            for ws, ilp, order, cut in ((0, 1, 0, 0), (0, 2, 1, 1), (1, 1, 1, 1), (1, 2, 0, 0), (1, 2, 1, 1), (1, 3, 1, 1), (1, 4, 0, 1)):
                if True:
                    ops.march_config(ws, ilp, order, cut)
                    d0, a0, _ = ops.shadow_march_fwd(depth, bits, P_L, samples=t, inside_bonus=5.0, want_argmin=True, variant=0)
                    assert torch.equal(d0, d1) and torch.equal(a0, a1), (ws, ilp, order)
                    o = ops.march_shade_fwd(albedo, depth, bits, P_L, amb, inside_bonus=5.0, samples=t, want=("rendered", "normals", "d_min"))
                    assert torch.equal(o["d_min"], d1), (ws, ilp)
                    if ref is None:
                        ref = o
                    assert torch.equal(o["rendered"], ref["rendered"]) and torch.equal(o["normals"], ref["normals"]), (ws, ilp)
    finally:
        ops.march_config(-1, 0, -1, -1)


def test_early_cutoff_is_exact_on_face_like_depth(ops):
    """Round 2: the default kernel stops a ray at the sample index beyond which the pixel -> light line is provably farther from
    every depth an in-mask sample can return than the ray's current minimum.  Smooth face-like depth (where the cut-off removes
    a third to two thirds of the in-mask samples), all 18 light directions plus lights below the image plane (descending
    lines use the minimum depth), background depth far above / below the face, depth of either sign, a nearly flat map:
    d_min and the arg-min must equal the literal kernel's bit for bit."""
    from geomconsistentfr_b200.synthetic import LIGHTS_18, synthetic_face
    H = W = 256
    dirs = list(LIGHTS_18) + [(0.6, 0.2, -0.5), (-0.3, -0.6, -0.2), (0.0, 0.0, 1.0), (0.02, -0.01, 0.9997)]
    base, m = synthetic_face(seed=3, noise=1.0)
    bg_hi = base.clone(); bg_hi[m == 0] = 400.0                    # background above the face: only the dilated-mask range keeps the cut-off useful
    bg_lo = base.clone(); bg_lo[m == 0] = -300.0
    bg_far = base.clone(); bg_far[m == 0] = 1.0e5                  # a pixel depth far outside the face's range (rounding-error bound of its rays)
    bg_neg = base.clone(); bg_neg[m == 0] = -1.0e5
    variants = [base, bg_hi, bg_lo, base - 80.0, 1e-3 * base, -base, bg_far, bg_neg]
    mask = (m * 255).view(1, H, W).cuda()
    bits = ops.mask_pack(mask)
    try:
        for vi, dm in enumerate(variants):
            for c0 in range(0, len(dirs), 8):
                L = torch.tensor(dirs[c0:c0 + 8], dtype=torch.float32)
                P_L = (4013.0 * torch.nn.functional.normalize(L, dim=1)).cuda()
                depth = dm.view(1, 1, H, W).repeat(L.shape[0], 1, 1, 1).cuda().contiguous()
                d1, a1, _ = ops.shadow_march_fwd(depth, bits, P_L, inside_bonus=5.0, want_argmin=True, variant=1)
                for cut in (1, 0):
                    ops.march_config(-1, 0, -1, cut)
                    d0, a0, _ = ops.shadow_march_fwd(depth, bits, P_L, inside_bonus=5.0, want_argmin=True, variant=0)
                    assert torch.equal(d0, d1), (vi, c0, cut, int((d0 != d1).sum()))
                    assert torch.equal(a0, a1), (vi, c0, cut, int((a0 != a1).sum()))
        # several lights per face and per-image masks go through the same range words
        ops.march_config(-1, 0, -1, 1)
        faces = [synthetic_face(seed=s, noise=2.0) for s in (1, 2)]
        depth = torch.stack([f[0] for f in faces]).view(2, 1, H, W).cuda()
        masks = torch.stack([f[1] for f in faces]).cuda()
        masks[1, :, :100] = 0
        bits2 = ops.mask_pack(masks)
        P_L = (4013.0 * torch.nn.functional.normalize(torch.tensor(dirs[2:8], dtype=torch.float32), dim=1)).cuda()      # 2 faces x 3 lights
        d0, a0, _ = ops.shadow_march_fwd(depth, bits2, P_L, want_argmin=True, variant=0)
        d1, a1, _ = ops.shadow_march_fwd(depth, bits2, P_L, want_argmin=True, variant=1)
        assert torch.equal(d0, d1) and torch.equal(a0, a1)
    finally:
        ops.march_config(-1, 0, -1, -1)


def test_shade_render_vs_oracle(ops, march):
    """TRAIN:353-369, 517-522 on the golden depth: normals, full/final shading, rendered.
    Tolerance 2e-5 absolute (all outputs are O(1); fp32 with a different summation order)."""
    B = 3
    depth = torch.from_numpy(march["depth"]).view(1, 1, 256, 256).repeat(B, 1, 1, 1)
    depth = depth + torch.arange(B).view(B, 1, 1, 1) * 7.0
    g = torch.Generator().manual_seed(7)
    albedo = torch.rand(B, 3, 256, 256, generator=g)
    d_min = 6.0 * torch.rand(B, 256, 256, generator=g)
    amb = torch.tensor([0.31, 0.45, 0.12])
    P_L = _light_pt([march["light_" + t] for t in ("right_above", "left_below", "inside")])
    n, _, amb_l, full = O.shade(depth, O.intrinsic_matrix(), P_L, amb)
    s = O.shadow_weight(d_min)
    fin, rend = O.render(albedo, s, full, amb_l)
    out = ops.shade_render_fwd(albedo.cuda(), depth.cuda(), d_min.cuda(), P_L.cuda(), amb.cuda())
    for k, ref in (("normals", n), ("full", full), ("final", fin), ("rendered", rend), ("shadow", s)):
        assert (out[k].cpu() - ref).abs().max() <= 2e-5, k


def test_ops_reject_cpu_tensors(ops):
    with pytest.raises(RuntimeError):
        ops.mask_pack(torch.zeros(1, 256, 256, dtype=torch.uint8))
    with pytest.raises(RuntimeError):
        ops.shadow_march_fwd(torch.zeros(1, 1, 256, 256), torch.zeros(1, 2048, dtype=torch.int32), torch.zeros(1, 3))


def test_fused_march_shade_equals_two_launches(ops, march):
    """gfr_march_shade_fwd (shading in the ray-march kernel's epilogue) == gfr_shadow_march_fwd + gfr_shade_render_fwd, bit
    for bit, also with several lights per face."""
    F_, L = 2, 3
    depth = torch.from_numpy(march["depth"]).view(1, 1, 256, 256).repeat(F_, 1, 1, 1).cuda()
    depth[1] += 5.0
    bits = ops.mask_pack(torch.from_numpy(march["mask_u8"]).view(1, 256, 256).cuda())
    P_L = _light_pt([march["light_" + t] for t in TAGS[:F_ * L]]).cuda()
    albedo = torch.rand(F_, 3, 256, 256, generator=torch.Generator().manual_seed(3)).cuda()
    amb = torch.tensor([0.3, 0.45]).cuda()
    d, _, _ = ops.shadow_march_fwd(depth, bits, P_L, inside_bonus=5.0)
    two = ops.shade_render_fwd(albedo, depth, d, P_L, amb)
    one = ops.march_shade_fwd(albedo, depth, bits, P_L, amb, inside_bonus=5.0, want=("shadow", "full", "final", "rendered", "normals", "d_min"))
    assert torch.equal(one["d_min"], d)
    for k in ("shadow", "full", "final", "rendered", "normals"):
        assert torch.equal(one[k], two[k]), k
