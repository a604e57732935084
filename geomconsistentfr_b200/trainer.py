"""The generator side of the reference's training step (TRAIN:606-656) on library kernels:

    forward (RelightNet in train mode) -> masked recon / depth / albedo losses (K5) + ambient L1 + light cosine +
    DSSIM (K4) -> backward (K2b, K1b, BN / conv backward) -> ONE flat-buffer gradient all-reduce -> fused Adam.

The PatchGAN discriminator (TRAIN:15-35) and its two loss terms (`0.01*BCE(D(composite), 1)` in the generator loss,
TRAIN:641-642, and the discriminator update, TRAIN:619-631) are NOT built yet: `GeneratorStep` trains the remaining six
terms of TRAIN:645 and reports which terms it used.  TRAIN = train_raytracing_relighting_CelebAHQ_DSSIM_8x.py."""
import torch

from .autograd import FlatAdam, MaskedLosses, dssim_loss

LOSS_TERMS = ("recon", "depth", "ambient", "lighting", "albedo", "DSSIM")     # TRAIN:672-682 names, minus the GAN terms


class GeneratorStep:
    def __init__(self, net, intrinsic_matrix, lr=None, group=None):
        self.net = net.train()
        self.K = intrinsic_matrix
        self.opt = FlatAdam(list(net.parameters()), lr=net.lr if lr is None else lr)      # TRAIN:589: Adam(lr=0.0001)
        self.group = group

    def losses(self, out, img, masks_fill, masks, depth_gt, albedo_gt, lighting_gt):
        """The reference's loss expressions (TRAIN:633-645).  img [B,H,W,3]; masks_fill / masks [B,H,W] in {0,1};
        depth_gt [B,H,W]; albedo_gt [B,H,W]; lighting_gt [B,4] = (ambient, lx, ly, lz)."""
        albedo, depth, _, _, _, rendered, unit_light, ambient_values = out
        B = img.shape[0]
        target = img.permute(0, 3, 1, 2).contiguous()
        mf, m = masks_fill.float().contiguous(), masks.float().contiguous()
        recon, depth_l, albedo_l = MaskedLosses.apply(rendered, depth[:, 0], albedo, target, depth_gt.float().contiguous(),
                                                      albedo_gt.float().contiguous(), mf, m)                    # TRAIN:633-639
        ambient_l = 2.5 * (ambient_values.reshape(B) - lighting_gt[:, 0]).abs().mean()                        # TRAIN:635
        light_l = torch.sum(1.0 - torch.sum(unit_light.reshape(B, 3) * lighting_gt[:, 1:4], dim=1)) / B        # TRAIN:636
        m3 = mf[:, None]
        composite = rendered * m3 + (1.0 - m3) * target
        dssim = dssim_loss(composite, target)                                                                  # TRAIN:643
        terms = dict(recon=recon, depth=depth_l, ambient=ambient_l, lighting=light_l, albedo=albedo_l, DSSIM=dssim)
        total = recon.float() + depth_l.float() + ambient_l + light_l + albedo_l.float() + dssim
        return total, terms

    # ---- the whole step as ONE CUDA graph (static shapes): ~630 kernel launches per step leave the CPU out of the loop
    def capture(self, img, epoch, masks_fill, masks, depth_gt, albedo_gt, lighting_gt, warmup=3):
        """Capture forward + losses + backward + gradient all-reduce + Adam for this batch shape; afterwards
        `step_graphed(...)` copies a batch into the static buffers and replays.  The Adam step counter lives on the
        device, BatchNorm buffers and parameters are updated in place by the replay."""
        self._static = [t.clone() for t in (img, masks_fill, masks, depth_gt, albedo_gt, lighting_gt)]
        self._epoch = epoch
        self._stream = torch.cuda.Stream(device=img.device)
        self._stream.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(self._stream):
            for _ in range(warmup):
                self.step(self._static[0], epoch, *self._static[1:])
            self._stream.synchronize()
            self._graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self._graph, stream=self._stream):
                self._static_out = self.step(self._static[0], epoch, *self._static[1:])
        torch.cuda.current_stream().wait_stream(self._stream)
        return self

    def step_graphed(self, img, masks_fill, masks, depth_gt, albedo_gt, lighting_gt):
        with torch.cuda.stream(self._stream):
            for dst, src in zip(self._static, (img, masks_fill, masks, depth_gt, albedo_gt, lighting_gt)):
                dst.copy_(src, non_blocking=True)
            self._graph.replay()
        return self._static_out

    def step(self, img, epoch, masks_fill, masks, depth_gt, albedo_gt, lighting_gt):
        """One optimiser step.  Returns (total, terms) as device tensors (no host sync)."""
        self.opt.zero_grad()                                                              # TRAIN:631
        B, H, W, _ = img.shape
        out = self.net(img, epoch, self.K, masks_fill.reshape(B, H, W, 1))                # TRAIN:618
        total, terms = self.losses(out, img, masks_fill, masks, depth_gt, albedo_gt, lighting_gt)
        total.backward()                                                                  # TRAIN:655
        scale = self.opt.all_reduce_grads(self.group)                                     # data parallel: one collective
        self.opt.step(grad_scale=scale)                                                   # TRAIN:656
        return total.detach(), {k: v.detach() for k, v in terms.items()}
