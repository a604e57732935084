"""GPU parity of the training-loss kernels (K4 SSIM fwd/bwd, K5 masked losses) and the fused Adam step, through the
autograd wrappers over the C ABI.  References: the oracle's pytorch_msssim restatement differentiated by autograd
(DSSIM parity is unpinned upstream), the reference's own loss expressions (TRAIN:633-639) evaluated in torch
float64, and torch.optim.Adam."""
import pytest
import torch

from oracle import relight_oracle as O

pytestmark = pytest.mark.gpu


def test_ssim_forward_backward_vs_oracle():
    from geomconsistentfr_b200 import SSIMPlanes, dssim_loss
    g = torch.Generator().manual_seed(2)
    N, C, H, W = 2, 3, 64, 80
    Y = torch.rand(N, C, H, W, generator=g)
    X = (Y + 0.2 * torch.randn(N, C, H, W, generator=g)).clamp(0, 1)
    Xr = X.clone().requires_grad_()
    loss_ref = 8.0 * (1 - O.ssim(Xr, Y, data_range=1.0, size_average=True, nonnegative_ssim=True)) / 2.0   # TRAIN:643
    loss_ref.backward()
    per_ref = O.ssim(X, Y, data_range=1.0, size_average=False).detach()        # mean over channels only
    Xc = X.clone().cuda().requires_grad_()
    per = SSIMPlanes.apply(Xc, Y.cuda(), 1.0)
    assert (per.mean(1).cpu() - per_ref).abs().max() <= 2e-6
    loss = dssim_loss(Xc, Y.cuda())
    assert abs(float(loss) - float(loss_ref)) <= 1e-5
    loss.backward()
    gref = Xr.grad
    assert float((Xc.grad.cpu() - gref).abs().max() / gref.abs().max()) <= 1e-4


def test_masked_losses_vs_reference_expressions():
    from geomconsistentfr_b200 import MaskedLosses
    g = torch.Generator().manual_seed(3)
    N, H, W = 3, 32, 48
    rendered, albedo = torch.rand(N, 3, H, W, generator=g), torch.rand(N, 3, H, W, generator=g)
    img = torch.rand(N, 3, H, W, generator=g)
    depth, depth_gt = torch.randn(N, 1, H, W, generator=g) * 50, torch.randn(N, H, W, generator=g) * 50
    albedo_gt = torch.rand(N, H, W, generator=g)
    mask_fill = (torch.rand(N, H, W, generator=g) > 0.4).double()
    mask = (torch.rand(N, H, W, generator=g) > 0.5).double() * mask_fill
    # the reference's expressions, float64 masks (TRAIN:633-639)
    r, d, a = (t.clone().requires_grad_() for t in (rendered, depth, albedo))
    m3 = mask_fill[:, None].repeat(1, 3, 1, 1)
    recon = 20.0 * ((r * m3 - img * m3) ** 2).sum() / m3.sum()
    dl = (d[:, 0] * mask - depth_gt * mask).abs().sum() / mask.sum()
    al = 5.0 * (a.mean(1) * mask_fill - albedo_gt * mask_fill).abs().sum() / mask_fill.sum()
    (recon + 2.0 * dl + 3.0 * al).backward()
    c = lambda t: t.float().cuda()
    rc, dc, ac = (c(t).requires_grad_() for t in (rendered, depth[:, 0], albedo))
    recon2, dl2, al2 = MaskedLosses.apply(rc, dc, ac, c(img), c(depth_gt), c(albedo_gt), c(mask_fill), c(mask))
    for x, y in ((recon2, recon), (dl2, dl), (al2, al)):
        assert abs(float(x) - float(y)) <= 1e-6 * max(1.0, abs(float(y)))
    (recon2 + 2.0 * dl2 + 3.0 * al2).backward()
    assert (rc.grad.cpu() - r.grad).abs().max() <= 1e-7 + 1e-5 * r.grad.abs().max()
    assert (dc.grad.cpu() - d.grad[:, 0]).abs().max() <= 1e-7 + 1e-5 * d.grad.abs().max()
    assert (ac.grad.cpu() - a.grad).abs().max() <= 1e-7 + 1e-5 * a.grad.abs().max()


def test_flat_adam_vs_torch_adam():
    from geomconsistentfr_b200 import FlatAdam
    g = torch.Generator(device="cuda").manual_seed(4)
    shapes = [(16, 3, 5, 5), (16,), (155, 64, 3, 3), (4, 128)]
    ref = [torch.randn(s, device="cuda", generator=g).requires_grad_() for s in shapes]
    mine = [p.detach().clone().requires_grad_() for p in ref]
    opt_ref = torch.optim.Adam(ref, lr=1e-4)                       # TRAIN:589
    opt = FlatAdam(mine, lr=1e-4)
    for _ in range(5):
        grads = [torch.randn(s, device="cuda", generator=g) for s in shapes]
        opt.zero_grad()
        for p, q, gr in zip(ref, mine, grads):
            p.grad = gr.clone()
            q.grad.copy_(gr)
        opt_ref.step()
        opt.step()
    for p, q in zip(ref, mine):
        assert (p - q).abs().max().item() <= 5e-7          # 1-2 ulp at |p| ~ 4 (different but equivalent op order)
