"""GPU parity of the backward kernels (K1b ray-march, K2b shade/render) through the autograd Functions, against the
CPU oracle differentiated by torch autograd (the reference has no hand-written backward: autograd through
TRAIN:353-369, 374-522 is what these kernels replace).  Tolerances are relative to the largest gradient magnitude:
fp32 chain rule vs autograd's fp32/fp64 mix agrees to ~1e-5; 2e-3 leaves room for atomicAdd ordering and arg-min ties."""
import numpy as np
import pytest
import torch

from oracle import relight_oracle as O

pytestmark = pytest.mark.gpu

LIGHTS = {
    "right_above": (0.6, 0.35, 0.7),        # x edge / y edge mix (TRAIN:445-458)
    "left": (-0.7574, 0.0, 0.6529),         # x edge only (TRAIN:399-403)
    "above": (0.0, 0.7071, 0.7071),         # y edge only (TRAIN:426-430)
    "below_left": (-0.6, -0.5, 0.62),       # both, lower-left corner (TRAIN:385-398)
    "inside": (0.004, 0.003, 1.0),          # the light projects inside the image: detached end point (TRAIN:423-425)
}


def _rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-12))


@pytest.mark.parametrize("tag", sorted(LIGHTS))
def test_march_backward_vs_oracle_autograd(tag):
    from geomconsistentfr_b200 import ShadowMarch, ops
    H = W = 64
    depth, mask = O.synthetic_face(seed=3, H=H, W=W, noise=2.0)
    depth = depth * 0.5
    g = torch.Generator().manual_seed(11)
    G = torch.randn(1, H, W, generator=g)
    L = torch.tensor([LIGHTS[tag]], dtype=torch.float32)
    P_L0 = O.light_point(L)[1]
    # oracle
    d_ref_in = depth.view(1, 1, H, W).clone().requires_grad_()
    P_ref = P_L0.clone().requires_grad_()
    d_ref, arg_ref = O.shadow_march(d_ref_in, mask.view(1, H, W), P_ref, return_argmin=True)
    (d_ref * G).sum().backward()
    # kernels
    d_in = depth.view(1, 1, H, W).cuda().requires_grad_()
    P_in = P_L0.cuda().requires_grad_()
    bits = ops.mask_pack(mask.view(1, H, W).cuda())
    d = ShadowMarch.apply(d_in, bits, P_in, 0.0)
    (d * G.cuda()).sum().backward()
    assert (d.cpu() - d_ref.detach()).abs().max() <= 2e-5 * max(1.0, float(d_ref[d_ref < 1e5].abs().max()))
    gd, gd_ref = d_in.grad.cpu()[0, 0], d_ref_in.grad[0, 0]
    scale = float(gd_ref.abs().max())
    bad = ((gd - gd_ref).abs() > 2e-3 * scale).float().mean().item()
    assert bad <= 2e-3, (tag, bad)                     # arg-min ties may route a handful of pixels differently
    assert float((gd - gd_ref).abs().sum() / gd_ref.abs().sum()) <= 2e-3
    gl, gl_ref = P_in.grad.cpu(), P_ref.grad
    assert _rel(gl, gl_ref) <= 2e-3, (tag, gl, gl_ref)


def test_shade_render_backward_vs_oracle_autograd():
    from geomconsistentfr_b200 import ShadeRender
    B, H, W = 2, 48, 64
    g = torch.Generator().manual_seed(5)
    depth = torch.stack([O.synthetic_face(seed=s, H=H, W=W, noise=1.5)[0] for s in (1, 2)])[:, None] * 0.6
    albedo = torch.rand(B, 3, H, W, generator=g)
    d_min = torch.rand(B, H, W, generator=g) * 4.0
    L = torch.tensor([(0.6, 0.35, 0.7), (-0.5, 0.2, 0.8)], dtype=torch.float32)
    P_L = O.light_point(L)[1]
    amb = torch.tensor([0.35, 0.5])
    Gs, Gf, Gfin = (torch.randn(B, H, W, generator=g) for _ in range(3))
    Gr, Gn = (torch.randn(B, 3, H, W, generator=g) for _ in range(2))
    K = O.intrinsic_matrix(H, W)

    ins = [t.clone().requires_grad_() for t in (albedo, depth, d_min, P_L, amb)]
    n, _, amb_l, full = O.shade(ins[1], K, ins[3], ins[4])
    s = O.shadow_weight(ins[2])
    final, rendered = O.render(ins[0], s, full, amb_l)
    ((rendered * Gr).sum() + (s * Gs).sum() + (full * Gf).sum() + (final * Gfin).sum() + (n * Gn).sum()).backward()

    cin = [t.clone().cuda().requires_grad_() for t in (albedo, depth, d_min, P_L, amb)]
    intr = (O.FOCAL, O.FOCAL, W / 2.0, H / 2.0, O.DEPTH_OFFSET, O.DIRECTIONAL_INTENSITY)
    s2, full2, final2, rendered2, n2 = ShadeRender.apply(*cin, intr)
    c = lambda t: t.cuda()
    ((rendered2 * c(Gr)).sum() + (s2 * c(Gs)).sum() + (full2 * c(Gf)).sum() + (final2 * c(Gfin)).sum() + (n2 * c(Gn)).sum()).backward()
    assert (rendered2.cpu() - rendered.detach()).abs().max() <= 2e-5
    for name, a, b in zip(("albedo", "depth", "d_min", "light", "ambient"), cin, ins):
        assert _rel(a.grad.cpu(), b.grad) <= 2e-3, (name, _rel(a.grad.cpu(), b.grad))


def test_render_chain_backward_vs_oracle_autograd():
    """depth/light/albedo/ambient -> march -> shade/render -> loss on rendered only, as in the training step
    (TRAIN:618, 633): gradients through K2b and K1b together."""
    from geomconsistentfr_b200 import ShadeRender, ShadowMarch, ops
    B, H, W = 2, 64, 64
    g = torch.Generator().manual_seed(8)
    faces = [O.synthetic_face(seed=s, H=H, W=W, noise=1.5) for s in (4, 5)]
    depth = torch.stack([f[0] for f in faces])[:, None] * 0.5
    masks = torch.stack([f[1] for f in faces])
    albedo = torch.rand(B, 3, H, W, generator=g)
    L = torch.tensor([(0.55, 0.3, 0.75), (-0.7, 0.1, 0.7)], dtype=torch.float32)
    P_L = O.light_point(L)[1]
    amb = torch.tensor([0.4, 0.3])
    Gr = torch.randn(B, 3, H, W, generator=g) * masks[:, None]
    K = O.intrinsic_matrix(H, W)

    ins = [t.clone().requires_grad_() for t in (albedo, depth, P_L, amb)]
    n, _, amb_l, full = O.shade(ins[1], K, ins[2], ins[3])
    d_ref = O.shadow_march(ins[1], masks, ins[2])
    _, rendered = O.render(ins[0], O.shadow_weight(d_ref), full, amb_l)
    (rendered * Gr).sum().backward()

    cin = [t.clone().cuda().requires_grad_() for t in (albedo, depth, P_L, amb)]
    bits = ops.mask_pack(masks.cuda())
    d = ShadowMarch.apply(cin[1], bits, cin[2], 0.0)
    intr = (O.FOCAL, O.FOCAL, W / 2.0, H / 2.0, O.DEPTH_OFFSET, O.DIRECTIONAL_INTENSITY)
    rendered2 = ShadeRender.apply(cin[0], cin[1], d, cin[2], cin[3], intr)[3]
    (rendered2 * Gr.cuda()).sum().backward()
    assert (rendered2.cpu() - rendered.detach()).abs().max() <= 5e-5
    for name, a, b in zip(("albedo", "depth", "light", "ambient"), cin, ins):
        ga, gb = a.grad.cpu(), b.grad
        if name == "depth":
            assert float((ga - gb).abs().sum() / gb.abs().sum()) <= 5e-3
        else:
            assert _rel(ga, gb) <= 5e-3, (name, ga, gb)
