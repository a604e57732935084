"""Differentiable operators of the relight path (operator level): `ShadowMarch` and `ShadeRender` are
torch.autograd.Functions whose forward AND backward are libgfr_b200 kernels (C ABI: gfr_shadow_march_fwd/_bwd,
gfr_shade_render_fwd/_bwd).  They replace what autograd records for TRAIN:353-369 and TRAIN:374-522
(TRAIN = train_raytracing_relighting_CelebAHQ_DSSIM_8x.py)."""
import ctypes

import numpy as np
import torch

from . import _lib, ops
from .ops import _need, _ptr, _stream


class ShadowMarch(torch.autograd.Function):
    """d_min = ShadowMarch.apply(depth[B,1,H,W], mask_bits[1|B,H*W/32] i32, light_pt[B,3], inside_bonus)
    Gradients: depth, light_pt.  (TRAIN:374-515 / TEST1:351-496)"""

    @staticmethod
    def forward(ctx, depth, mask_bits, light_pt, inside_bonus=0.0):
        depth = depth.contiguous()
        light_pt = light_pt.contiguous()
        d_min, argmin, _ = ops.shadow_march_fwd(depth, mask_bits, light_pt, inside_bonus=inside_bonus, want_argmin=True)
        ctx.save_for_backward(depth, light_pt, argmin)
        return d_min

    @staticmethod
    def backward(ctx, g_dmin):
        depth, light_pt, argmin = ctx.saved_tensors
        B, _, H, W = depth.shape
        g_dmin = _need(g_dmin, torch.float32, "g_dmin")
        g_depth = torch.zeros_like(depth)
        g_light = torch.zeros_like(light_pt)
        t = ops.reference_samples()
        rc = _lib.load().gfr_shadow_march_bwd(_ptr(depth), _ptr(light_pt), _ptr(argmin), _ptr(g_dmin),
                                              t.ctypes.data_as(ctypes.c_void_p), int(t.shape[0]), _ptr(g_depth), _ptr(g_light),
                                              B, H, W, _stream())
        _lib.check(rc, "gfr_shadow_march_bwd"); ops._count()
        return g_depth, None, g_light, None


class ShadeRender(torch.autograd.Function):
    """shadow, full, final, rendered, normals = ShadeRender.apply(albedo, depth, d_min, light_pt, ambient, intr)
    intr = (fx, fy, cx, cy, depth_offset, intensity) python floats.  Gradients: albedo, depth, d_min, light_pt, ambient.
    (TRAIN:353-369, 517-522)"""

    @staticmethod
    def forward(ctx, albedo, depth, d_min, light_pt, ambient, intr):
        albedo, depth, d_min = albedo.contiguous(), depth.contiguous(), d_min.contiguous()
        light_pt, ambient = light_pt.contiguous(), ambient.reshape(-1).contiguous()
        fx, fy, cx, cy, off, inten = intr
        o = ops.shade_render_fwd(albedo, depth, d_min, light_pt, ambient, fx, fy, cx, cy, off, inten)
        ctx.save_for_backward(albedo, depth, d_min, light_pt, ambient)
        ctx.intr = np.array(intr, dtype=np.float32)
        return o["shadow"], o["full"], o["final"], o["rendered"], o["normals"]

    @staticmethod
    def backward(ctx, g_shadow, g_full, g_final, g_rendered, g_normals):
        albedo, depth, d_min, light_pt, ambient = ctx.saved_tensors
        B, _, H, W = depth.shape
        c = lambda g: None if g is None else _need(g, torch.float32, "grad")
        g_shadow, g_full, g_final, g_rendered, g_normals = c(g_shadow), c(g_full), c(g_final), c(g_rendered), c(g_normals)
        g_albedo = torch.empty_like(albedo)
        g_depth = torch.zeros_like(depth)
        g_dmin = torch.empty_like(d_min)
        g_light = torch.zeros_like(light_pt)
        g_amb = torch.zeros_like(ambient)
        rc = _lib.load().gfr_shade_render_bwd(
            _ptr(albedo), _ptr(depth), _ptr(d_min), _ptr(light_pt), _ptr(ambient), ctx.intr.ctypes.data_as(ctypes.c_void_p),
            _ptr(g_shadow), _ptr(g_full), _ptr(g_final), _ptr(g_rendered), _ptr(g_normals), _ptr(g_albedo), _ptr(g_depth),
            _ptr(g_dmin), _ptr(g_light), _ptr(g_amb), B, H, W, _stream())
        _lib.check(rc, "gfr_shade_render_bwd"); ops._count()
        return g_albedo, g_depth, g_dmin, g_light, g_amb, None


class SSIMPlanes(torch.autograd.Function):
    """per = SSIMPlanes.apply(X[N,C,H,W], Y[N,C,H,W], data_range) -> [N,C] mean SSIM of every plane (the `.mean(-1)` of
    pytorch_msssim._ssim, TRAIN:643).  Gradient w.r.t. X only (Y is the target image)."""

    @staticmethod
    def forward(ctx, X, Y, data_range=1.0):
        X, Y = _need(X, torch.float32, "X"), _need(Y, torch.float32, "Y")
        N, C, H, W = X.shape
        P = N * C
        sums = torch.zeros(P, dtype=torch.float64, device=X.device)
        gm = torch.empty((3, P, H - 10, W - 10), dtype=torch.float32, device=X.device) if X.requires_grad else None
        rc = _lib.load().gfr_ssim_fwd(_ptr(X), _ptr(Y), _ptr(sums), _ptr(gm), P, H, W, float(data_range), _stream())
        _lib.check(rc, "gfr_ssim_fwd"); ops._count()
        ctx.save_for_backward(X, Y, gm)
        return (sums / float((H - 10) * (W - 10))).to(torch.float32).view(N, C)

    @staticmethod
    def backward(ctx, g_per):
        X, Y, gm = ctx.saved_tensors
        N, C, H, W = X.shape
        scale = (g_per.reshape(-1).to(torch.float32) / float((H - 10) * (W - 10))).contiguous()
        gX = torch.empty_like(X)
        rc = _lib.load().gfr_ssim_bwd(_ptr(X), _ptr(Y), _ptr(gm), _ptr(scale), _ptr(gX), N * C, H, W, _stream())
        _lib.check(rc, "gfr_ssim_bwd"); ops._count()
        return gX, None, None


def dssim_loss(composite, target, weight=8.0):
    """TRAIN:643: weight * (1 - ssim(composite, target, data_range=1, size_average=True, nonnegative_ssim=True)) / 2."""
    per = SSIMPlanes.apply(composite, target, 1.0)
    return weight * (1.0 - torch.relu(per).mean()) / 2.0


class MaskedLosses(torch.autograd.Function):
    """recon, depth_l, albedo_l = MaskedLosses.apply(rendered, depth, albedo, img_nchw, depth_gt, albedo_gt, mask_fill, mask)
    — the three masked terms of TRAIN:633-639 (fp64 sums like the reference) with their gradients from one kernel pass."""

    @staticmethod
    def forward(ctx, rendered, depth, albedo, img_nchw, depth_gt, albedo_gt, mask_fill, mask):
        f = lambda t, n: _need(t, torch.float32, n)
        rendered, depth, albedo = f(rendered, "rendered"), f(depth, "depth"), f(albedo, "albedo")
        img_nchw, depth_gt, albedo_gt = f(img_nchw, "img"), f(depth_gt, "depth_gt"), f(albedo_gt, "albedo_gt")
        mask_fill, mask = f(mask_fill, "mask_fill"), f(mask, "mask")
        N, _, H, W = rendered.shape
        sums = torch.empty(5, dtype=torch.float64, device=rendered.device)
        g_r, g_d, g_a = torch.empty_like(rendered), torch.empty_like(depth), torch.empty_like(albedo)
        rc = _lib.load().gfr_masked_losses(_ptr(rendered), _ptr(img_nchw), _ptr(depth), _ptr(depth_gt), _ptr(albedo),
                                           _ptr(albedo_gt), _ptr(mask_fill), _ptr(mask), _ptr(sums), _ptr(g_r), _ptr(g_d),
                                           _ptr(g_a), N, H, W, _stream())
        _lib.check(rc, "gfr_masked_losses"); ops._count(2)
        ctx.save_for_backward(g_r, g_d, g_a)
        return 20.0 * sums[2] / (3.0 * sums[0]), sums[3] / sums[1], 5.0 * sums[4] / sums[0]

    @staticmethod
    def backward(ctx, g_recon, g_depth_l, g_albedo_l):
        g_r, g_d, g_a = ctx.saved_tensors
        return (g_r * g_recon.to(torch.float32), g_d * g_depth_l.to(torch.float32), g_a * g_albedo_l.to(torch.float32),
                None, None, None, None, None)


class FlatAdam:
    """torch.optim.Adam(lr, betas=(0.9, 0.999), eps=1e-8) over ONE flat fp32 buffer that the parameters are views of
    (TRAIN:589-590, 656) — a single fused kernel per step; `grad_scale` folds the 1/world_size of the data-parallel
    gradient all-reduce.

    Like torch.optim.Adam the state is PER PARAMETER: every parameter tensor is a segment with its own step count and bias
    correction (on the device, so a step replays from a CUDA graph), and a parameter that received no gradient is skipped
    entirely — torch skips `p.grad is None` parameters, which is what the epoch-gated skip blocks are until their gate
    opens (TRAIN:245,258,271,283); they then start at step 1 with a properly bias-corrected first update.  Here every
    `.grad` is a view of the flat gradient buffer, so "has a gradient" is declared by the owner: `set_active(flags)`."""

    def __init__(self, params, lr=1e-4, betas=(0.9, 0.999), eps=1e-8):
        self.params = [p for p in params]
        dev = self.params[0].device
        n = sum(p.numel() for p in self.params)
        self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(n, dtype=torch.float32, device=dev)
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        o, starts = 0, []
        with torch.no_grad():
            for p in self.params:                      # re-seat every parameter (and its .grad) as a view of the flat buffers
                k = p.numel()
                starts.append(o)
                self.flat[o:o + k].copy_(p.reshape(-1))
                p.data = self.flat[o:o + k].view_as(p)
                p.grad = self.grad[o:o + k].view_as(p)
                o += k
        self.lr, self.betas, self.eps = lr, betas, eps
        self.seg_start = torch.tensor(starts + [n], dtype=torch.int64, device=dev)
        # per segment {step, 1-b1^t, sqrt(1-b2^t), active}, advanced on the device
        self.seg_state = torch.zeros((len(self.params), 4), dtype=torch.float32, device=dev)
        self.seg_state[:, 3] = 1.0

    def zero_grad(self):
        self.grad.zero_()

    def set_active(self, flags):
        """flags[i] False: parameter i gets no gradient in the coming steps (torch: `.grad is None`) — it is not updated and
        its step count does not advance.  A plain device copy: call it between (not inside) graph replays."""
        f = torch.as_tensor([1.0 if a else 0.0 for a in flags], dtype=torch.float32)
        if f.numel() != len(self.params):
            raise ValueError("set_active needs one flag per parameter")
        self.seg_state[:, 3].copy_(f.to(self.seg_state.device))

    # ---- checkpointing, in torch.optim.Adam's own state_dict layout (per-parameter step / exp_avg / exp_avg_sq; parameters
    # that never had a gradient have no entry), so a checkpoint written here resumes under the reference's
    # `torch.optim.Adam` (TRAIN:589-590) and vice versa
    def state_dict(self):
        steps = self.seg_state[:, 0].detach().cpu().tolist()
        st, o = {}, 0
        for i, p in enumerate(self.params):
            k = p.numel()
            if steps[i] > 0:
                st[i] = {"step": torch.tensor(float(steps[i])), "exp_avg": self.exp_avg[o:o + k].view_as(p).detach().cpu().clone(),
                         "exp_avg_sq": self.exp_avg_sq[o:o + k].view_as(p).detach().cpu().clone()}
            o += k
        group = {"lr": self.lr, "betas": tuple(self.betas), "eps": self.eps, "weight_decay": 0, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "decoupled_weight_decay": False, "params": list(range(len(self.params)))}
        return {"state": st, "param_groups": [group]}

    def load_state_dict(self, sd):
        g = sd["param_groups"][0]
        if len(g["params"]) != len(self.params):
            raise ValueError("optimizer state holds %d parameters, this optimiser %d" % (len(g["params"]), len(self.params)))
        self.lr, self.betas, self.eps = float(g["lr"]), tuple(g["betas"]), float(g["eps"])
        state, o = torch.zeros((len(self.params), 3), dtype=torch.float64), 0
        with torch.no_grad():
            for i, p in enumerate(self.params):
                k = p.numel()
                e = sd["state"].get(i)
                if e is None:                      # never stepped (torch keeps no state for it)
                    self.exp_avg[o:o + k].zero_(); self.exp_avg_sq[o:o + k].zero_()
                else:
                    self.exp_avg[o:o + k].copy_(e["exp_avg"].reshape(-1)); self.exp_avg_sq[o:o + k].copy_(e["exp_avg_sq"].reshape(-1))
                    t = float(e["step"])
                    state[i] = torch.tensor([t, 1.0 - self.betas[0] ** t, (1.0 - self.betas[1] ** t) ** 0.5], dtype=torch.float64)
                o += k
            self.seg_state[:, :3].copy_(state.float())

    def all_reduce_grads(self, group=None):
        """ONE collective per optimiser step over the flat gradient buffer (sum); returns the 1/world factor."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
            return 1.0
        ev = getattr(self, "ar_events", None)         # optional (start, end) CUDA events (external=True under graph capture)
        if ev is not None:
            ev[0].record()
        dist.all_reduce(self.grad, op=dist.ReduceOp.SUM, group=group)
        if ev is not None:
            ev[1].record()
        return 1.0 / dist.get_world_size(group)

    def step(self, grad_scale=1.0):
        rc = _lib.load().gfr_adam_step_segments(_ptr(self.flat), _ptr(self.grad), _ptr(self.exp_avg), _ptr(self.exp_avg_sq),
                                                self.flat.numel(), _ptr(self.seg_start), _ptr(self.seg_state), len(self.params),
                                                self.lr, self.betas[0], self.betas[1], self.eps, float(grad_scale), _stream())
        _lib.check(rc, "gfr_adam_step_segments"); ops._count(2)
        ops.bump_param_generation()
