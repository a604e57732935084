"""GPU parity of the TRAIN-mode path (batch-statistics BatchNorm, autograd through every unit) against what the
UNMODIFIED reference produced in train() mode for B = 3 (tests/golden/train.npz, made by oracle/make_golden.py:
TRAIN forward + loss.backward() with loss = masked L2 recon + DSSIM, TRAIN:633,643), plus unit-level gradient checks
against torch autograd (cuDNN fp32)."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _rel(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-20))


@pytest.mark.parametrize("deconv,cin,cout,S,res,post", [
    (False, 16, 16, 32, True, True), (True, 32, 16, 32, False, False), (False, 64, 155, 16, True, False),
    (True, 128, 64, 16, False, False),
])
def test_conv_bn_act_unit_vs_torch_autograd(deconv, cin, cout, S, res, post):
    from geomconsistentfr_b200 import ops, train_ops as T
    g = torch.Generator(device="cuda").manual_seed(cin + cout)
    N = 3
    mod = (torch.nn.ConvTranspose2d if deconv else torch.nn.Conv2d)(cin, cout, 3, padding=1).cuda()
    bn = torch.nn.BatchNorm2d(cout).cuda()
    with torch.no_grad():
        bn.weight.uniform_(0.5, 1.5, generator=g); bn.bias.normal_(generator=g)
    ctot = cin + (27 if cin == 128 else 0)                      # exercise "leading channels of a wider tensor"
    x = torch.randn(N, ctot, S, S, device="cuda", generator=g)
    r = torch.randn(N, cout, S, S, device="cuda", generator=g) if res else None
    p = torch.randn(N, cout, S // 2, S // 2, device="cuda", generator=g) if post else None
    Gy = torch.randn(N, cout, S, S, device="cuda", generator=g)
    # torch reference in float64 (cuDNN's fp32 backward-filter algorithms are neither deterministic nor equally accurate)
    mod64 = (torch.nn.ConvTranspose2d if deconv else torch.nn.Conv2d)(cin, cout, 3, padding=1).cuda().double()
    bn64 = torch.nn.BatchNorm2d(cout).cuda().double()
    with torch.no_grad():
        mod64.weight.copy_(mod.weight); mod64.bias.copy_(mod.bias); bn64.weight.copy_(bn.weight); bn64.bias.copy_(bn.bias)
    xr = x.double().requires_grad_()
    rr = r.double().requires_grad_() if res else None
    pr = p.double().requires_grad_() if post else None
    pre = bn64(mod64(xr[:, :cin]))
    y_ref = F.leaky_relu(pre + rr if res else pre, 0.2)
    if post:
        y_ref = y_ref + F.interpolate(pr, scale_factor=2, mode="nearest")
    (y_ref * Gy.double()).sum().backward()
    ref = dict(w=mod64.weight.grad.float(), b=mod64.bias.grad.float(), gamma=bn64.weight.grad.float(), beta=bn64.bias.grad.float(),
               x=xr.grad.float())
    y_ref = y_ref.float()
    bn = bn64
    rr_grad = rr.grad.float() if res else None
    pr_grad = pr.grad.float() if post else None
    bn2 = torch.nn.BatchNorm2d(cout).cuda()
    # ours
    xc = ops.nchw_to_c4(x).data.requires_grad_()
    rc = ops.nchw_to_c4(r).data.requires_grad_() if res else None
    pc = ops.nchw_to_c4(p).data.requires_grad_() if post else None
    meta = dict(cin=cin, cout=cout, deconv=deconv, act=1, post_shift=1 if post else 0, bn=bn2)
    bn2.weight.data.copy_(bn.weight.data); bn2.bias.data.copy_(bn.bias.data)
    y = T.ConvBNAct.apply(xc, mod.weight, mod.bias, bn2.weight, bn2.bias, rc, pc, meta)
    got_y = ops.c4_to_nchw(ops.C4(y.detach(), cout))
    assert (got_y - y_ref.detach()).abs().max() <= 2e-4
    assert (bn2.running_mean - bn.running_mean.float()).abs().max() <= 1e-5 and (bn2.running_var - bn.running_var.float()).abs().max() <= 1e-4
    (y * ops.nchw_to_c4(Gy).data).sum().backward()
    assert _rel(mod.weight.grad, ref["w"]) <= 2e-3
    assert _rel(bn2.weight.grad, ref["gamma"]) <= 2e-3 and _rel(bn2.bias.grad, ref["beta"]) <= 2e-3
    assert mod.bias.grad.abs().max() <= 1e-3 * max(1.0, float(ref["w"].abs().max()))      # exactly 0 in exact arithmetic (BN removes the mean)
    assert _rel(ops.c4_to_nchw(ops.C4(xc.grad, ctot)), ref["x"]) <= 2e-3
    if res:
        assert _rel(ops.c4_to_nchw(ops.C4(rc.grad, cout)), rr_grad) <= 1e-4
    if post:
        assert _rel(ops.c4_to_nchw(ops.C4(pc.grad, cout)), pr_grad) <= 1e-4


@pytest.fixture(scope="module")
def train_net():
    from geomconsistentfr_b200 import RelightNet
    n = RelightNet(batch_size=3)
    n.load_state_dict(torch.load(os.path.join(G, "model_epoch99.pth"), map_location="cpu"), strict=True)
    return n.float().cuda().train()


def test_train_mode_forward_and_gradients_vs_reference(train_net):
    """Forward (train-mode BN) and the gradients the reference's autograd produced, for the reference's own B = 3."""
    from geomconsistentfr_b200 import dssim_loss, intrinsic_matrix
    t = np.load(os.path.join(G, "train.npz"))
    f = np.load(os.path.join(G, "ffhq.npz"))
    x = torch.from_numpy(f["q"][t["sel"]] / 1020.0).float().cuda()
    mt = torch.from_numpy(t["masks01"].astype(np.float64)).view(3, 256, 256, 1).cuda()
    train_net.zero_grad()
    out = train_net(x, 200, intrinsic_matrix().cuda(), mt)
    assert len(out) == 8
    albedo, depth, shadow, amb_l, full, rendered, unit_l, amb_v = out
    depth.retain_grad()
    tol = dict(albedo=5e-5, depth=2e-2, shadow=5e-3, rendered=5e-3)
    for k, v in (("albedo", albedo), ("depth", depth), ("shadow", shadow), ("rendered", rendered)):
        d = np.abs(v.detach().cpu().numpy() - t[k]).max()
        assert d <= tol[k], (k, d)
    assert np.abs(unit_l.detach().cpu().numpy().reshape(3, 3) - t["unit_light"]).max() <= 2e-5
    assert np.abs(amb_v.detach().cpu().numpy().reshape(3) - t["ambient"]).max() <= 2e-5
    m3 = mt.permute(0, 3, 1, 2).repeat(1, 3, 1, 1).float()
    target = x.permute(0, 3, 1, 2).contiguous()
    comp = rendered * m3 + (1.0 - m3) * target
    loss = 20.0 * torch.sum((rendered * m3 - target * m3) ** 2) / torch.sum(m3) + dssim_loss(comp, target)   # TRAIN:633,643
    assert abs(float(loss) - float(t["loss"])) <= 2e-4 * float(t["loss"])
    loss.backward()
    gd, gd_ref = depth.grad.cpu().numpy(), t["grad_depth"]
    assert np.abs(gd - gd_ref).sum() / np.abs(gd_ref).sum() <= 2e-2
    for name, got in (("grad_sl2_w", train_net.linear_SL2.weight.grad), ("grad_sl2_b", train_net.linear_SL2.bias.grad),
                      ("grad_depth_head_w", train_net.conv_depth_c2_o.weight.grad),
                      ("grad_albedo_head_w", train_net.conv_albedo_c2_o.weight.grad)):
        ref = t[name]
        err = np.abs(got.cpu().numpy() - ref).max() / np.abs(ref).max()
        assert err <= 2e-2, (name, err)
    # every parameter received a finite gradient
    for n_, p_ in train_net.named_parameters():
        assert p_.grad is not None and torch.isfinite(p_.grad).all(), n_


def test_generator_step_losses_and_all_gradients_vs_oracle():
    """One full generator step body (TRAIN:618, 633-645 without the two PatchGAN terms, 655) on the library kernels vs the
    CPU oracle in train() mode differentiated by torch autograd: every loss term and EVERY parameter gradient."""
    from geomconsistentfr_b200 import RelightNet, intrinsic_matrix
    from geomconsistentfr_b200.trainer import GeneratorStep
    from oracle import relight_oracle as O
    f = np.load(os.path.join(G, "ffhq.npz"))
    sel = [0, 4]
    B = len(sel)
    sd = torch.load(os.path.join(G, "model_epoch99.pth"), map_location="cpu")
    img = torch.from_numpy(f["q"][sel] / 1020.0).float()
    mf = torch.from_numpy((f["masks"][sel] > 128).astype(np.float32))
    m = mf.clone(); m[:, 100:140, 90:170] = 0                                # the depth mask excludes nose/mouth (TRAIN:610 vs 612)
    gen = torch.Generator().manual_seed(1)
    depth_gt = torch.stack([O.synthetic_face(seed=s)[0] for s in (1, 2)]) * 0.5
    albedo_gt = torch.rand(B, 256, 256, generator=gen)
    light_gt = torch.tensor([[0.45, 0.5145, 0.0, 0.8575], [0.55, -0.5843, 0.0, 0.8115]])

    # ---- oracle (CPU, train mode)
    ref = O.RelightNetOracle(); ref.load_state_dict(sd); ref.train()
    out = ref.forward_train(img, 200, O.intrinsic_matrix(), mf.double().view(B, 256, 256, 1))
    albedo, depth, _, _, _, rendered, unit_l, amb_v = out
    m3 = mf[:, None].repeat(1, 3, 1, 1).double(); target = img.permute(0, 3, 1, 2)
    terms_ref = dict(
        recon=20.0 * ((rendered * m3 - target * m3) ** 2).sum() / m3.sum(),
        depth=(depth[:, 0] * m.double() - depth_gt * m.double()).abs().sum() / m.double().sum(),
        ambient=2.5 * (amb_v.reshape(B) - light_gt[:, 0]).abs().mean(),
        lighting=torch.sum(1 - torch.sum(unit_l.reshape(B, 3) * light_gt[:, 1:4], dim=1)) / B,
        albedo=5.0 * (albedo.mean(1) * mf.double() - albedo_gt * mf.double()).abs().sum() / mf.double().sum(),
        DSSIM=8.0 * (1 - O.ssim((rendered * m3 + (1 - m3) * target).float(), target, data_range=1.0, size_average=True,
                                nonnegative_ssim=True)) / 2.0)
    sum(terms_ref.values()).backward()

    # ---- library
    net = RelightNet(batch_size=B); net.load_state_dict(sd, strict=True); net = net.float().cuda().train()
    step = GeneratorStep(net, intrinsic_matrix().cuda())
    c = lambda t: t.cuda()
    step.opt.zero_grad()
    o = net(c(img), 200, step.K, c(mf).view(B, 256, 256, 1))
    total, terms = step.losses(o, c(img), c(mf), c(m), c(depth_gt), c(albedo_gt), c(light_gt))
    for k in terms_ref:
        a, b = float(terms[k]), float(terms_ref[k])
        assert abs(a - b) <= 2e-3 * max(abs(b), 1e-3), (k, a, b)
    total.backward()
    worst, l1 = {}, {}
    for (n1, p1), (n2, p2) in zip(net.named_parameters(), ref.named_parameters()):
        assert n1 == n2
        if n1.endswith(".bias") and ("conv" in n1 or "deconv" in n1) and not n1.endswith("c2_o.bias"):
            continue                                  # conv bias before a train-mode BN: gradient is 0 up to rounding on both sides
        g1, g2 = p1.grad.cpu(), p2.grad
        worst[n1] = float((g1 - g2).abs().max() / (g2.abs().max() + 1e-12))
        l1[n1] = float((g1 - g2).abs().sum() / (g2.abs().sum() + 1e-12))
    # Two metrics per parameter tensor.  L1-relative (the bulk): what 3xTF32 convs + fp64 BN sums + atomics deliver.  Max-norm
    # (the worst element): carries the isolated discontinuities of the path — a LeakyReLU / BN unit whose pre-activation is
    # within rounding of 0 flips its slope (1 vs 0.2), a ray-march arg-min tie moves the gradient to another depth pixel, a
    # max-pool tie to another input — each moves one contribution by O(1) of its size in the oracle's own fp32 run as well.
    top = sorted(worst.items(), key=lambda kv: -kv[1])[:6]
    print("worst max-norm gradient errors:", ["%s %.1e (L1 %.1e)" % (k, v, l1[k]) for k, v in top])
    print("worst L1 gradient errors:", ["%s %.1e" % kv for kv in sorted(l1.items(), key=lambda kv: -kv[1])[:6]],
          "median max-norm %.1e, median L1 %.1e" % (np.median(list(worst.values())), np.median(list(l1.values()))))
    bad = {k: v for k, v in worst.items() if v > 3e-2}
    assert not bad, bad
    assert np.median(list(worst.values())) <= 5e-3
    assert max(l1.values()) <= 1e-2, max(l1.values())
    assert np.median(list(l1.values())) <= 2e-3


def test_patchgan_forward_backward_vs_oracle():
    """PatchGAN (TRAIN:15-35) in train() mode on the library kernels (space-to-depth + tcgen05 3x3, conv5 on CUDA cores) vs
    the torch oracle: logits, input gradient (what flows back into the generator, TRAIN:641-642) and all parameter gradients."""
    from geomconsistentfr_b200 import PatchGAN
    from oracle import relight_oracle as O
    torch.manual_seed(0)
    ref = O.PatchGANOracle().cuda().train()
    mine = PatchGAN().cuda().train()
    mine.load_state_dict(ref.state_dict(), strict=True)
    g = torch.Generator(device="cuda").manual_seed(9)
    img = torch.rand(4, 3, 256, 256, device="cuda", generator=g)
    Gl = torch.randn(4, 1, 15, 15, device="cuda", generator=g)
    xr = img.clone().requires_grad_()
    lr_ = ref(xr)
    (lr_ * Gl).sum().backward()
    xm = img.clone().requires_grad_()
    lm = mine(xm)
    assert lm.shape == (4, 1, 15, 15)
    assert _rel(lm.detach(), lr_.detach()) <= 2e-4
    (lm * Gl).sum().backward()
    assert _rel(xm.grad, xr.grad) <= 5e-3
    for (n1, p1), (n2, p2) in zip(mine.named_parameters(), ref.named_parameters()):
        assert n1 == n2
        if n1 in ("conv2.bias", "conv3.bias", "conv4.bias"):
            continue                                   # bias before a train-mode BN: zero gradient up to rounding
        assert _rel(p1.grad, p2.grad) <= 1e-2, (n1, _rel(p1.grad, p2.grad))
    for b1, b2 in zip(mine.buffers(), ref.buffers()):
        assert (b1.float() - b2.float()).abs().max() <= 1e-4


def test_full_train_iteration_losses_vs_oracle():
    """One iteration of TRAIN:617-656 (j = 0: the discriminator is updated before the generator loss is evaluated):
    all ten loss terms vs the CPU oracle driven exactly like the reference (torch.optim.Adam on the oracle PatchGAN)."""
    import torch.nn.functional as F
    from geomconsistentfr_b200 import PatchGAN, RelightNet, intrinsic_matrix
    from geomconsistentfr_b200.trainer import TrainStep
    from oracle import relight_oracle as O
    f = np.load(os.path.join(G, "ffhq.npz"))
    sel = [3, 6]
    B = len(sel)
    sd = torch.load(os.path.join(G, "model_epoch99.pth"), map_location="cpu")
    img = torch.from_numpy(f["q"][sel] / 1020.0).float()
    mf = torch.from_numpy((f["masks"][sel] > 128).astype(np.float32))
    depth_gt = torch.stack([O.synthetic_face(seed=s)[0] for s in (3, 4)]) * 0.5
    albedo_gt = torch.rand(B, 256, 256, generator=torch.Generator().manual_seed(2))
    light_gt = torch.tensor([[0.45, 0.5145, 0.0, 0.8575], [0.55, -0.5843, 0.0, 0.8115]])
    torch.manual_seed(5)
    d_ref = O.PatchGANOracle().train()
    d_sd = {k: v.clone() for k, v in d_ref.state_dict().items()}

    # ---- oracle, the reference's order of operations
    ref = O.RelightNetOracle(); ref.load_state_dict(sd); ref.train()
    opt_d = torch.optim.Adam(d_ref.parameters(), lr=1e-4)
    opt_d.zero_grad()
    out = ref.forward_train(img, 200, O.intrinsic_matrix(), mf.double().view(B, 256, 256, 1))
    albedo, depth, _, _, _, rendered, unit_l, amb_v = out
    m3 = mf[:, None].repeat(1, 3, 1, 1); target = img.permute(0, 3, 1, 2)
    comp = rendered * m3 + (1.0 - m3) * target
    lf, lr_ = d_ref(comp), d_ref(target)
    t_ref = dict(discriminator_fake=0.01 * F.binary_cross_entropy_with_logits(lf, torch.zeros_like(lf)),
                 discriminator_real=0.01 * F.binary_cross_entropy_with_logits(lr_, torch.ones_like(lr_)))
    (t_ref["discriminator_fake"] + t_ref["discriminator_real"]).backward(retain_graph=True)
    opt_d.step()
    lf2 = d_ref(comp)
    t_ref["generator"] = 0.01 * F.binary_cross_entropy_with_logits(lf2, torch.ones_like(lf2))
    t_ref["recon"] = 20.0 * ((rendered * m3.double() - target * m3.double()) ** 2).sum() / m3.double().sum()
    t_ref["DSSIM"] = 8.0 * (1 - O.ssim(comp, target, data_range=1.0, size_average=True, nonnegative_ssim=True)) / 2.0

    # ---- library
    net = RelightNet(batch_size=B); net.load_state_dict(sd, strict=True); net = net.float().cuda().train()
    D = PatchGAN(); D.load_state_dict(d_sd, strict=True); D = D.cuda().train()
    step = TrainStep(net, D, intrinsic_matrix().cuda())
    c = lambda t: t.cuda()
    total, terms = step.step(c(img), 200, c(mf), c(mf), c(depth_gt), c(albedo_gt), c(light_gt), j=0)
    assert set(terms) == {"recon", "depth", "ambient", "lighting", "albedo", "DSSIM", "generator", "discriminator",
                          "discriminator_real", "discriminator_fake"}                       # TRAIN:672-682
    for k, v in t_ref.items():
        a, b = float(terms[k]), float(v)
        assert abs(a - b) <= 2e-3 * max(abs(b), 1e-4), (k, a, b)
    # the discriminator really moved (Adam step on every parameter that has a gradient) and stayed finite
    moved = sum(float((p.detach().cpu() - d_sd[n]).abs().max()) > 0 for n, p in D.named_parameters())
    assert moved >= 10 and all(torch.isfinite(p).all() for p in D.parameters())
    assert torch.isfinite(total)


def test_patchgan_bf16_mode_vs_oracle():
    """BASELINE configs[2] "bf16 CNN": the same PatchGAN with bf16 conv operands (fp32 accumulation in TMEM, fp32 BatchNorm /
    activations): logits and gradients agree with the fp32 oracle to bf16 accuracy (2^-8 per operand)."""
    from geomconsistentfr_b200 import PatchGAN
    from oracle import relight_oracle as O
    torch.manual_seed(0)
    ref = O.PatchGANOracle().cuda().train()
    mine = PatchGAN().cuda().train()
    mine.train_precision = 4
    mine.load_state_dict(ref.state_dict(), strict=True)
    g = torch.Generator(device="cuda").manual_seed(9)
    img = torch.rand(4, 3, 256, 256, device="cuda", generator=g)
    Gl = torch.randn(4, 1, 15, 15, device="cuda", generator=g)
    xr = img.clone().requires_grad_()
    lr_ = ref(xr)
    (lr_ * Gl).sum().backward()
    xm = img.clone().requires_grad_()
    lm = mine(xm)
    assert _rel(lm.detach(), lr_.detach()) <= 3e-2
    (lm * Gl).sum().backward()
    l1 = lambda a, b: float((a - b).abs().sum() / b.abs().sum())
    assert l1(xm.grad, xr.grad) <= 0.15                  # measured 0.086: four bf16 layers forward and backward (3xTF32: 2e-4)
    for (n1, p1), (n2, p2) in zip(mine.named_parameters(), ref.named_parameters()):
        if n1 in ("conv2.bias", "conv3.bias", "conv4.bias"):
            continue
        assert l1(p1.grad, p2.grad) <= 0.15, (n1, l1(p1.grad, p2.grad))


@pytest.mark.parametrize("prec,tol", [(3, 2e-3), (4, 4e-2)])
def test_conv_bn_act_unit_precisions_vs_torch_autograd(prec, tol):
    """A RelightNet train unit (3x3) in 3xTF32 and in bf16 against torch fp64 autograd (L1-relative gradients)."""
    from geomconsistentfr_b200 import ops, train_ops as T
    g = torch.Generator(device="cuda").manual_seed(prec)
    N, cin, cout, S = 3, 32, 64, 32
    mod = torch.nn.Conv2d(cin, cout, 3, padding=1).cuda().double()
    bn = torch.nn.BatchNorm2d(cout).cuda().double()
    x = torch.randn(N, cin, S, S, device="cuda", generator=g)
    Gy = torch.randn(N, cout, S, S, device="cuda", generator=g)
    xr = x.double().requires_grad_()
    y_ref = F.leaky_relu(bn(mod(xr)), 0.2)
    (y_ref * Gy.double()).sum().backward()
    bn2 = torch.nn.BatchNorm2d(cout).cuda()
    w, b = mod.weight.detach().float().requires_grad_(), mod.bias.detach().float().requires_grad_()
    xc = ops.nchw_to_c4(x).data.requires_grad_()
    meta = dict(cin=cin, cout=cout, deconv=False, act=1, post_shift=0, bn=bn2, precision=prec)
    y = T.ConvBNAct.apply(xc, w, b, bn2.weight, bn2.bias, None, None, meta)
    l1 = lambda a, r: float((a.double() - r).abs().sum() / r.abs().sum())
    assert l1(ops.c4_to_nchw(ops.C4(y.detach(), cout)), y_ref.detach()) <= tol
    (y * ops.nchw_to_c4(Gy).data).sum().backward()
    assert l1(w.grad, mod.weight.grad) <= tol and l1(ops.c4_to_nchw(ops.C4(xc.grad, cin)), xr.grad) <= tol
    assert l1(bn2.weight.grad, bn.weight.grad) <= tol


@pytest.mark.parametrize("cin,cout,deconv,S", [(16, 16, False, 64), (16, 32, False, 32), (16, 16, True, 48), (13, 16, False, 40), (32, 16, False, 32)])
def test_wgrad_tc_bf16_both_mma_arrangements_vs_torch(cin, cout, deconv, S):
    """gfr_conv_wgrad_tc_bf16 (weight gradient on tcgen05, bf16 operands) against torch's conv weight gradient in fp64 on the
    bf16-rounded operands, in both MMA arrangements (one MMA per tap / the pixel-shift form for <= 16 input channels): fp32
    accumulation order is the only difference -> 2e-5 of the largest gradient."""
    from geomconsistentfr_b200 import _lib, ops, train_ops as T
    g = torch.Generator(device="cuda").manual_seed(cin * 100 + cout + S)
    N = 3
    x = torch.randn(N, cin, S, S, device="cuda", generator=g).bfloat16().float()
    gy = torch.randn(N, cout, S, S, device="cuda", generator=g).bfloat16().float()
    if deconv:       # ConvTranspose2d(k=3, s=1, p=1): weight [cin, cout, 3, 3]
        w = torch.zeros(cin, cout, 3, 3, device="cuda", dtype=torch.float64, requires_grad=True)
        F.conv_transpose2d(x.double(), w, padding=1).mul(gy.double()).sum().backward()
    else:
        w = torch.zeros(cout, cin, 3, 3, device="cuda", dtype=torch.float64, requires_grad=True)
        F.conv2d(x.double(), w, padding=1).mul(gy.double()).sum().backward()
    xc, gc = ops.nchw_to_c4(x).data, ops.nchw_to_c4(gy).data
    try:
        for form in (0, 1):
            _lib.check(_lib.load().gfr_wgrad_tc_config(form), "gfr_wgrad_tc_config")
            wp = torch.zeros(w.shape, device="cuda", dtype=torch.float32, requires_grad=True)
            wp.grad = torch.zeros_like(wp)
            T._wgrad(xc, gc, wp, None, deconv, cin, cout, taps=9, precision=4)
            err = float((wp.grad.double() - w.grad).abs().max() / w.grad.abs().max())
            assert err <= 2e-5, (form, err)
    finally:
        _lib.load().gfr_wgrad_tc_config(-1)


@pytest.mark.gpu
@pytest.mark.parametrize("C,S,nchw", [(3, 32, True), (8, 16, False), (64, 32, False), (6, 16, False)])
def test_space_to_depth_pad_and_its_backward_vs_pixel_unshuffle(C, S, nchw):
    """PatchGAN's input re-blocking (csrc/patchgan.cu): out[n, 4c + 2dy + dx, yo, xo] = pad1(x)[n, c, 2yo + dy, 2xo + dx] and its
    adjoint — the vectorised C4 kernels (C % 4 == 0), the scalar C4 kernels (C = 6) and the planar-input form, bit-exact
    against F.pixel_unshuffle of the padded tensor (a pure re-indexing)."""
    from geomconsistentfr_b200 import ops
    from geomconsistentfr_b200.patchgan import _SpaceToDepthPad
    torch.manual_seed(C)
    x = torch.randn(2, C, S, S, device="cuda")
    xr = x.clone().requires_grad_()
    ref = F.pixel_unshuffle(F.pad(xr, (1, 1, 1, 1)), 2)
    Gy = torch.randn_like(ref)
    ref.backward(Gy)
    xin = x.clone().requires_grad_() if nchw else ops.nchw_to_c4(x).data.requires_grad_()
    out = _SpaceToDepthPad.apply(xin, C, nchw)
    assert torch.equal(ops.c4_to_nchw(ops.C4(out.detach(), 4 * C)), ref.detach())
    (out * ops.nchw_to_c4(Gy).data).sum().backward()
    got = xin.grad if nchw else ops.c4_to_nchw(ops.C4(xin.grad, C))
    assert torch.equal(got, xr.grad)
