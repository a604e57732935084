"""Test helper: drives a (RelightNet-like, PatchGAN-like) pair through iterations of the reference's training loop
TRAIN:606-656 with two torch.optim.Adam — the oracle side of the trajectory tests.  The loop body is the reference's order of
operations (D loss on (composite, real), D step iff j % 5 == 0, the seven generator terms with D(fake) re-evaluated after the
D step, G step); the fixtures it is compared with come from the reference's UNMODIFIED main() (oracle/make_golden_train_iter.py)."""
import torch
import torch.nn.functional as F

from oracle import relight_oracle as O

NAMES = ("total", "recon", "depth", "ambient", "lighting", "albedo", "generator", "discriminator", "discriminator_real",
         "discriminator_fake", "DSSIM")


def oracle_iteration(net, D, opt, opt_d, batch, epoch, j, K):
    """One iteration on the oracle modules.  batch = (img [B,H,W,3], masks_fill [B,H,W], masks [B,H,W], depth_gt [B,H,W],
    albedo_gt [B,H,W], lightings [B,4]) float32 host tensors.  Returns the 11 loss terms as floats."""
    img, mf, m, depth_gt, albedo_gt, light = batch
    B, H, W, _ = img.shape
    opt_d.zero_grad()                                                                   # TRAIN:617
    out = net.forward_train(img, epoch, K, mf.double().view(B, H, W, 1))               # TRAIN:618
    albedo, depth, _, _, _, rendered, unit_l, amb_v = out
    m3 = mf[:, None].repeat(1, 3, 1, 1)
    target = img.permute(0, 3, 1, 2)
    comp = rendered * m3 + (1.0 - m3) * target
    lf, lr_ = D(comp), D(target)                                                        # TRAIN:619-620
    d_fake = 0.01 * F.binary_cross_entropy_with_logits(lf, torch.zeros_like(lf))
    d_real = 0.01 * F.binary_cross_entropy_with_logits(lr_, torch.ones_like(lr_))
    d_loss = d_fake + d_real
    if j % 5 == 0:                                                                      # TRAIN:624-626
        d_loss.backward(retain_graph=True)
        opt_d.step()
    opt.zero_grad()                                                                     # TRAIN:631
    m3d, md, mfd = m3.double(), m.double(), mf.double()
    recon = 20.0 * ((rendered * m3d - target * m3d) ** 2).sum() / m3d.sum()             # TRAIN:633
    depth_l = (depth[:, 0] * md - depth_gt * md).abs().sum() / md.sum()                 # TRAIN:634
    ambient = 2.5 * (amb_v.reshape(B) - light[:, 0]).abs().mean()                       # TRAIN:635
    lighting = torch.sum(1 - torch.sum(unit_l.reshape(B, 3) * light[:, 1:4], dim=1)) / B   # TRAIN:636
    albedo_l = 5.0 * (albedo.mean(1) * mfd - albedo_gt * mfd).abs().sum() / mfd.sum()   # TRAIN:637-639
    lf2 = D(comp)                                                                       # TRAIN:641
    g_loss = 0.01 * F.binary_cross_entropy_with_logits(lf2, torch.ones_like(lf2))
    dssim = 8.0 * (1 - O.ssim(comp, target, data_range=1.0, size_average=True, nonnegative_ssim=True)) / 2.0   # TRAIN:643
    total = recon + depth_l + ambient + lighting + albedo_l + g_loss + dssim             # TRAIN:645
    total.backward()                                                                    # TRAIN:655
    opt.step()                                                                          # TRAIN:656
    vals = (total, recon, depth_l, ambient, lighting, albedo_l, g_loss, d_loss, d_real, d_fake, dssim)
    return [float(v) for v in vals]
