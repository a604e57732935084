"""The generator side of the reference's training step (TRAIN:606-656) on library kernels:

    forward (RelightNet in train mode) -> masked recon / depth / albedo losses (K5) + ambient L1 + light cosine +
    DSSIM (K4) -> backward (K2b, K1b, BN / conv backward) -> ONE flat-buffer gradient all-reduce -> fused Adam.

`GeneratorStep` trains the six non-adversarial terms of TRAIN:645; `TrainStep` is the reference's full iteration
(TRAIN:617-656): discriminator loss on (composite, real) with an Adam step every GD_ratio-th iteration, then the
generator loss with the 0.01*BCE(D(composite), 1) term (TRAIN:641-642) and its Adam step.
TRAIN = train_raytracing_relighting_CelebAHQ_DSSIM_8x.py."""
import contextlib
import os

import torch
import torch.nn.functional as F

from . import ops, train_ops
from .autograd import FlatAdam, MaskedLosses, dssim_loss

LOSS_TERMS = ("recon", "depth", "ambient", "lighting", "albedo", "DSSIM")     # TRAIN:672-682 names, minus the GAN terms


class GeneratorStep:
    def __init__(self, net, intrinsic_matrix, lr=None, group=None):
        self.net = net.train()
        self.K = intrinsic_matrix
        self.opt = FlatAdam(list(net.parameters()), lr=net.lr if lr is None else lr)      # TRAIN:589: Adam(lr=0.0001)
        self.group = group
        self._active_sig = None
        self.batch_packs = os.environ.get("GFR_TRAIN_PACK_PLAN", "1") != "0"      # A/B switch: 0 = every layer packs its own operand
        self.bn_arena = train_ops.ZeroArena()
        self.use_bn_arena = os.environ.get("GFR_TRAIN_BN_ARENA", "1") != "0"      # A/B switch: 0 = a memset node per BatchNorm pass
        self.pack_plan = train_ops.PackPlan(list(net.parameters()))      # the generator's packed conv operands: one launch per step

    def _declare_active(self, epoch):
        """Parameters of encoder-skip blocks whose gate is closed get no gradient: tell the optimiser (torch.optim.Adam skips
        `.grad is None`).  A host->device copy, so it only happens when the gate signature changes (never inside a capture:
        the capture's eager warm-up steps have set it)."""
        flags = self.net.active_parameter_flags(epoch)
        sig = tuple(flags)
        if sig != self._active_sig:
            self.opt.set_active(flags)
            self._active_sig = sig

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
        ops.bump_param_generation()                 # the replay wrote parameters and BN buffers behind torch's back
        return self._static_out

    def step(self, img, epoch, masks_fill, masks, depth_gt, albedo_gt, lighting_gt, **kw):
        """One optimiser step.  Returns (total, terms) as device tensors (no host sync)."""
        with self.pack_plan, (self.bn_arena if self.use_bn_arena else contextlib.nullcontext()):
            if self.use_bn_arena:
                self.bn_arena.begin(img.device)        # every BatchNorm pass's fp64 scratch: one memset
            if self.batch_packs:
                self.pack_plan.run()                       # every conv operand of the generator from the current parameters, one launch
            try:
                return self._step(img, epoch, masks_fill, masks, depth_gt, albedo_gt, lighting_gt, **kw)
            finally:
                self.pack_plan.invalidate()            # the optimiser moved the parameters (TRAIN:656)

    def _step(self, img, epoch, masks_fill, masks, depth_gt, albedo_gt, lighting_gt):
        self._declare_active(epoch)
        self.opt.zero_grad()                                                              # TRAIN:631
        B, H, W, _ = img.shape
        out = self.net(img, epoch, self.K, masks_fill.reshape(B, H, W, 1))                # TRAIN:618
        total, terms = self.losses(out, img, masks_fill, masks, depth_gt, albedo_gt, lighting_gt)
        total.backward()                                                                  # TRAIN:655
        scale = self.opt.all_reduce_grads(self.group)                                     # data parallel: one collective
        self.opt.step(grad_scale=scale)                                                   # TRAIN:656
        return total.detach(), {k: v.detach() for k, v in terms.items()}


class TrainStep(GeneratorStep):
    """The reference's training iteration with the PatchGAN (TRAIN:617-656).  Two equivalences are used to skip work
    whose results the reference throws away: (1) `d_loss.backward(retain_graph=True)` (TRAIN:625) also back-propagates
    into the generator, but `optimizer.zero_grad()` (TRAIN:631) clears those gradients before they are used — here the
    discriminator sees `composite.detach()`; (2) `total_loss.backward()` (TRAIN:655) also fills the discriminator's
    parameter gradients, which `optimizer_patchgan.zero_grad()` (TRAIN:617) discards — here the discriminator's weights
    are frozen for that pass (its input gradient still flows).  Parameter updates are identical."""

    def __init__(self, net, patchgan, intrinsic_matrix, lr=None, group=None):
        super().__init__(net, intrinsic_matrix, lr, group)
        self.D = patchgan.train()
        self.D.train_precision = getattr(net, "train_precision", 3)          # one operand precision for both networks
        self.opt_d = FlatAdam(list(patchgan.parameters()), lr=net.lr if lr is None else lr)      # TRAIN:590
        self.GD_ratio = net.GD_ratio
        self.dedup_d = os.environ.get("GFR_TRAIN_DEDUP_D", "1") != "0"      # one discriminator pass over the composite instead of two on iterations without a D update (A/B switch)

    def _d_frozen(self, composite, record=False):
        """D(composite) with the discriminator's weights frozen (its input gradient still flows)."""
        for p in self.D.parameters():
            p.requires_grad_(False)
        try:
            if record:
                with train_ops.record_running_updates() as rec:
                    out = self.D(composite)
                self._fake_records = rec.records
                return out
            return self.D(composite)
        finally:
            for p in self.D.parameters():
                p.requires_grad_(True)

    def _step(self, img, epoch, masks_fill, masks, depth_gt, albedo_gt, lighting_gt, j=0):
        """Iteration j of an epoch (`step(..., j=j)`).  Returns (total, terms) with the reference's loss names (TRAIN:672-682)."""
        B, H, W, _ = img.shape
        update_d = (j % self.GD_ratio) == 0                                               # TRAIN:624
        self._declare_active(epoch)
        self.opt_d.zero_grad()                                                            # TRAIN:617
        out = self.net(img, epoch, self.K, masks_fill.reshape(B, H, W, 1))                # TRAIN:618
        rendered = out[5]
        target = img.permute(0, 3, 1, 2).contiguous()
        m3 = masks_fill.float()[:, None]
        composite = rendered * m3 + (1.0 - m3) * target
        fake_records = None
        if update_d or not self.dedup_d:
            logits_fake = self.D(composite.detach())                                      # TRAIN:619
        else:
            # No discriminator update this iteration (TRAIN:624): D(composite) of TRAIN:641 is the same function of the same
            # input as D(composite) of TRAIN:619, so ONE pass serves both — run here (the reference's BatchNorm update order is
            # fake, real, fake) with frozen weights and the input gradient the generator loss needs; the running-statistics
            # update of the pass that is not run is replayed from this pass's batch sums after D(real).
            logits_fake2 = self._d_frozen(composite, record=True)
            fake_records, logits_fake = self._fake_records, logits_fake2.detach()
        if update_d:
            logits_real = self.D(target)                                                  # TRAIN:620
        else:
            with torch.no_grad():                                                         # (d_loss is only reported on these iterations)
                logits_real = self.D(target)
        if fake_records is not None:
            train_ops._BN.replay_running_updates(fake_records)                            # the buffers' third update (TRAIN:641)
        d_fake = 0.01 * F.binary_cross_entropy_with_logits(logits_fake, torch.zeros_like(logits_fake))   # TRAIN:621
        d_real = 0.01 * F.binary_cross_entropy_with_logits(logits_real, torch.ones_like(logits_real))    # TRAIN:622
        d_loss = d_fake + d_real
        if update_d:
            d_loss.backward()                                                             # TRAIN:625
            self.opt_d.step(grad_scale=self.opt_d.all_reduce_grads(self.group))           # TRAIN:626
        self.opt.zero_grad()                                                              # TRAIN:631
        total, terms = self.losses(out, img, masks_fill, masks, depth_gt, albedo_gt, lighting_gt)
        if fake_records is None:
            logits_fake2 = self._d_frozen(composite)                                      # TRAIN:641 (after the D update)
        g_loss = 0.01 * F.binary_cross_entropy_with_logits(logits_fake2, torch.ones_like(logits_fake2))  # TRAIN:642
        total = total + g_loss
        total.backward()                                                                  # TRAIN:655
        self.opt.step(grad_scale=self.opt.all_reduce_grads(self.group))                   # TRAIN:656
        terms = dict(terms, generator=g_loss, discriminator=d_loss, discriminator_real=d_real, discriminator_fake=d_fake)
        return total.detach(), {k: v.detach() for k, v in terms.items()}

    # two graphs: iterations that update the discriminator and iterations that do not
    def capture(self, img, epoch, masks_fill, masks, depth_gt, albedo_gt, lighting_gt, warmup=2):
        self._static = [t.clone() for t in (img, masks_fill, masks, depth_gt, albedo_gt, lighting_gt)]
        self._stream = torch.cuda.Stream(device=img.device)
        self._stream.wait_stream(torch.cuda.current_stream())
        self._graphs, self._outs = {}, {}
        with torch.cuda.stream(self._stream):
            for _ in range(warmup):
                for j in (0, 1):
                    self.step(self._static[0], epoch, *self._static[1:], j=j)
            self._stream.synchronize()
            for j in (0, 1):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self._stream):
                    self._outs[j] = self.step(self._static[0], epoch, *self._static[1:], j=j)
                self._graphs[j] = g
        torch.cuda.current_stream().wait_stream(self._stream)
        return self

    def step_graphed(self, img, masks_fill, masks, depth_gt, albedo_gt, lighting_gt, j=0):
        k = 0 if (j % self.GD_ratio) == 0 else 1
        with torch.cuda.stream(self._stream):
            for dst, src in zip(self._static, (img, masks_fill, masks, depth_gt, albedo_gt, lighting_gt)):
                dst.copy_(src, non_blocking=True)
            self._graphs[k].replay()
        ops.bump_param_generation()
        return self._outs[k]
