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
