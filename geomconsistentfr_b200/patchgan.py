"""Drop-in `PatchGAN` (TRAIN:15-35; TRAIN = train_raytracing_relighting_CelebAHQ_DSSIM_8x.py): same attribute /
parameter names (conv1..conv5, bn2..bn4), forward(img[B,3,256,256]) -> logits [B,1,15,15], train mode (batch-statistics
BatchNorm, autograd) like the reference uses it (TRAIN:564-568, 619-620, 641).

conv1..conv4 (4x4, stride 2, pad 1) run on the tcgen05 3x3 kernel over a space-to-depth of their input; the 3x3 operand
is a linear re-indexing of the 4x4 parameter (`_w4x4_to_s2d3x3`, differentiable glue on the weight tensor), so the
unit's weight gradient flows back to the parameter through autograd.  conv5 (4x4, stride 1, 512 -> 1) is a small
CUDA-core kernel."""
import torch
import torch.nn as nn

from . import _lib, ops, train_ops
from .ops import _need, _ptr, _stream
from .train_ops import _chk


def _w4x4_to_s2d3x3(w):
    """w [Cout,Cin,4,4] -> W' [Cout,4*Cin,3,3] with W'[co, 4c+2dy+dx, a, b] = w[co, c, ky(a,dy), kx(b,dx)], where input row
    2y+ky-1 = (tap a, phase dy): ky 0 -> (0,1), 1 -> (1,0), 2 -> (1,1), 3 -> (2,0); the other two (a,dy) pairs are zero."""
    Cout, Cin = w.shape[0], w.shape[1]
    out = w.new_zeros((Cout, Cin, 2, 2, 3, 3))                       # [co, c, dy, dx, a, b]
    amap = ((0, 1), (1, 0), (1, 1), (2, 0))                          # ky -> (a, dy)
    for ky, (a, dy) in enumerate(amap):
        for kx, (b, dx) in enumerate(amap):
            out[:, :, dy, dx, a, b] = w[:, :, ky, kx]
    return out.reshape(Cout, Cin * 4, 3, 3)


class _SpaceToDepth(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, C, nchw):
        if nchw:
            N, _, H, W = x.shape
        else:
            N, _, H, W, _ = x.shape
        out = torch.empty((N, C, H // 2, W // 2, 4), dtype=torch.float32, device=x.device)
        _chk(_lib.load().gfr_space_to_depth(_ptr(x.contiguous()), _ptr(out), N, C, H, W, int(nchw), _stream()), "gfr_space_to_depth")
        ctx.cfg = (x.shape, C, H, W, nchw)
        return out

    @staticmethod
    def backward(ctx, g):
        shape, C, H, W, nchw = ctx.cfg
        g_x = torch.zeros(shape, dtype=torch.float32, device=g.device)
        _chk(_lib.load().gfr_depth_to_space(_ptr(g.contiguous()), _ptr(g_x), shape[0], C, H, W, int(nchw), _stream()), "gfr_depth_to_space")
        return g_x, None, None


class _ConvLReLU(torch.autograd.Function):
    """y = LeakyReLU(conv3x3(x) + b) without BatchNorm (conv1, TRAIN:29): fused epilogue forward, tcgen05 dgrad, CUDA-core wgrad."""

    @staticmethod
    def forward(ctx, x, w, b, cin, cout):
        NT = train_ops._nt_for(cout)
        N, G, H, W, _ = x.shape
        y = torch.empty((N, (cout + 3) // 4, H, W, 4), dtype=torch.float32, device=x.device)
        packed = train_ops._pack_dev(w, False, False, cin, cout, NT)
        _chk(_lib.load().gfr_conv3x3_tc_fwd(_ptr(x), _ptr(packed), _ptr(b), None, None, _ptr(y), N, cin, G, cout, H, W, NT, 0, 1, 1.0,
                                            3, 0, 1.0, 1.0, _stream()), "gfr_conv3x3_tc_fwd")
        ctx.save_for_backward(x, w, y)
        ctx.cfg = (cin, cout)
        return y

    @staticmethod
    def backward(ctx, g_y):
        x, w, y = ctx.saved_tensors
        cin, cout = ctx.cfg
        N, G, H, W, _ = x.shape
        g_raw = torch.empty_like(y)
        _chk(_lib.load().gfr_lrelu_bwd_c4(_ptr(y), _ptr(g_y.contiguous()), _ptr(g_raw), y.numel(), _stream()), "gfr_lrelu_bwd_c4")
        g_x = None
        if ctx.needs_input_grad[0]:
            NTd = train_ops._nt_for(cin)
            zero_b = torch.zeros(cin, dtype=torch.float32, device=x.device)
            g_x = train_ops._conv_raw(g_raw, cout, train_ops._pack_dev(w, False, True, cin, cout, NTd), zero_b, cin, NTd)
        g_w = g_b = None
        if ctx.needs_input_grad[1]:
            g_w = torch.zeros_like(w)
            g_b4 = torch.zeros(((cout + 3) // 4) * 4, dtype=torch.float32, device=x.device)
            _chk(_lib.load().gfr_conv3x3_wgrad(_ptr(x), _ptr(g_raw), _ptr(g_w), _ptr(g_b4), 0, N, cin, G, cout, H, W, _stream()),
                 "gfr_conv3x3_wgrad", 2)
            g_b = g_b4[:cout].contiguous()
        return g_x, g_w, g_b, None, None


class _Conv5(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w, b):
        N, G, H, W, _ = x.shape
        out = torch.empty((N, 1, H - 1, W - 1), dtype=torch.float32, device=x.device)
        _chk(_lib.load().gfr_conv4x4s1_to1_fwd(_ptr(x), _ptr(w), _ptr(b), _ptr(out), N, G * 4, H, W, _stream()), "gfr_conv4x4s1_to1_fwd")
        ctx.save_for_backward(x, w)
        return out

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        N, G, H, W, _ = x.shape
        g_x = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        g_w = torch.zeros_like(w) if ctx.needs_input_grad[1] else None
        g_b = torch.zeros(1, dtype=torch.float32, device=x.device) if g_w is not None else None
        n = (1 if g_x is not None else 0) + (1 if g_w is not None else 0)
        _chk(_lib.load().gfr_conv4x4s1_to1_bwd(_ptr(x), _ptr(w), _ptr(g.contiguous()), _ptr(g_x), _ptr(g_w), _ptr(g_b), N, G * 4, H, W,
                                               _stream()), "gfr_conv4x4s1_to1_bwd", n)
        return g_x, g_w, g_b


class PatchGAN(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 4, stride=2, padding=(1, 1))            # TRAIN:18-27
        self.conv2 = nn.Conv2d(64, 128, 4, stride=2, padding=(1, 1))
        self.conv3 = nn.Conv2d(128, 256, 4, stride=2, padding=(1, 1))
        self.conv4 = nn.Conv2d(256, 512, 4, stride=2, padding=(1, 1))
        self.conv5 = nn.Conv2d(512, 1, 4, stride=1, padding=(1, 1))
        self.bn2, self.bn3, self.bn4 = nn.BatchNorm2d(128), nn.BatchNorm2d(256), nn.BatchNorm2d(512)

    def forward(self, img):
        """img [B,3,H,W] fp32 CUDA (NCHW) -> logits [B,1,H/16-1,W/16-1] (TRAIN:28-35)."""
        if not img.is_cuda:
            raise RuntimeError("PatchGAN (geomconsistentfr_b200) runs on CUDA only")
        if not self.training:
            raise NotImplementedError("the reference only ever runs the discriminator in train() mode (TRAIN:564-568)")
        x = _SpaceToDepth.apply(img.float(), 3, True)                                     # [B,12,128,128]
        x = _ConvLReLU.apply(x, _w4x4_to_s2d3x3(self.conv1.weight), self.conv1.bias, 12, 64)
        for conv, bn, cin, cout in ((self.conv2, self.bn2, 64, 128), (self.conv3, self.bn3, 128, 256), (self.conv4, self.bn4, 256, 512)):
            x = _SpaceToDepth.apply(x, cin, False)
            meta = dict(cin=4 * cin, cout=cout, deconv=False, act=1, post_shift=0, bn=bn)
            x = train_ops.ConvBNAct.apply(x, _w4x4_to_s2d3x3(conv.weight), conv.bias, bn.weight, bn.bias, None, None, meta)
        return _Conv5.apply(x, self.conv5.weight, self.conv5.bias)
