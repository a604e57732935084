"""Drop-in `PatchGAN` (TRAIN:15-35; TRAIN = train_raytracing_relighting_CelebAHQ_DSSIM_8x.py): same attribute /
parameter names (conv1..conv5, bn2..bn4), forward(img[B,3,256,256]) -> logits [B,1,15,15], train mode (batch-statistics
BatchNorm, autograd) like the reference uses it (TRAIN:564-568, 619-620, 641).

conv1..conv4 (4x4, stride 2, pad 1) run on the tcgen05 conv kernel as 2x2-TAP convolutions over the space-to-depth of their
1-padded input (no structurally-zero weights; round 1 used a 3x3 form that multiplied 5/9 zeros); the 2x2 operand is a
permute of the 4x4 parameter (`_w4x4_to_taps2x2`, differentiable glue on the weight tensor), so the unit's weight
gradient flows back to the parameter through autograd.  conv5 (4x4, stride 1, 512 -> 1) is a small
CUDA-core kernel."""
import torch
import torch.nn as nn

from . import _lib, ops, train_ops
from .ops import _need, _ptr, _stream
from .train_ops import _chk


def _w4x4_to_taps2x2(w):
    """w [Cout,Cin,4,4] -> W' [Cout,4*Cin,2,2] with W'[co, 4c+2dy+dx, a, b] = w[co, c, 2a+dy, 2b+dx]: the 4x4 / stride 2 / pad 1
    convolution as a 2x2-tap / stride 1 convolution over the space-to-depth of the 1-padded input (block rows y, y+1 hold the
    input rows 2y-1 .. 2y+2).  A pure re-indexing (one permute), so the weight gradient flows back through autograd."""
    Cout, Cin = w.shape[0], w.shape[1]
    return w.view(Cout, Cin, 2, 2, 2, 2).permute(0, 1, 3, 5, 2, 4).reshape(Cout, Cin * 4, 2, 2)


class _SpaceToDepthPad(torch.autograd.Function):
    """[N,C,H,W] (NCHW planes or C4) -> C4 [N,4C,H/2+1,W/2+1]: space-to-depth of the 1-padded input (csrc/patchgan.cu)."""

    @staticmethod
    def forward(ctx, x, C, nchw):
        if nchw:
            N, _, H, W = x.shape
        else:
            N, _, H, W, _ = x.shape
        out = torch.empty((N, C, H // 2 + 1, W // 2 + 1, 4), dtype=torch.float32, device=x.device)
        _chk(_lib.load().gfr_space_to_depth_pad(_ptr(x.contiguous()), _ptr(out), N, C, H, W, int(nchw), _stream()), "gfr_space_to_depth_pad")
        ctx.cfg = (x.shape, C, H, W, nchw)
        return out

    @staticmethod
    def backward(ctx, g):
        shape, C, H, W, nchw = ctx.cfg
        g_x = torch.zeros(shape, dtype=torch.float32, device=g.device) if (not nchw and C % 4) else torch.empty(shape, dtype=torch.float32, device=g.device)
        _chk(_lib.load().gfr_depth_to_space_pad(_ptr(g.contiguous()), _ptr(g_x), shape[0], C, H, W, int(nchw), _stream()), "gfr_depth_to_space_pad")
        return g_x, None, None


class _ConvLReLU(torch.autograd.Function):
    """y = LeakyReLU(conv2x2taps(x) + b) without BatchNorm (conv1, TRAIN:29): fused epilogue forward, tcgen05 dgrad, CUDA-core wgrad."""

    @staticmethod
    def forward(ctx, x, w, b, cin, cout, prec):
        NT = train_ops._nt_for(cout, 4, prec)
        packed = train_ops._pack_dev(w, False, False, cin, cout, NT, 4, prec)
        y = train_ops._conv_raw(x, cin, packed, b, cout, NT, 4, 0, prec, act=1)
        ctx.save_for_backward(x, w, y, b)
        ctx.cfg = (cin, cout, prec)
        return y

    @staticmethod
    def backward(ctx, g_y):
        x, w, y, b = ctx.saved_tensors
        cin, cout, prec = ctx.cfg
        g_raw = torch.empty_like(y)
        _chk(_lib.load().gfr_lrelu_bwd_c4(_ptr(y), _ptr(g_y.contiguous()), _ptr(g_raw), y.numel(), _stream()), "gfr_lrelu_bwd_c4")
        g_x = None
        if ctx.needs_input_grad[0]:
            NTd = train_ops._nt_for(cin, 4, prec)
            zero_b = torch.zeros(cin, dtype=torch.float32, device=x.device)
            g_x = train_ops._conv_raw(g_raw, cout, train_ops._pack_dev(w, False, True, cin, cout, NTd, 4, prec), zero_b, cin, NTd, 4, 1, prec)
        g_w = g_b = None
        if ctx.needs_input_grad[1]:
            g_w, g_b = train_ops._wgrad(x, g_raw, w, b, False, cin, cout, 4, prec)
        return g_x, g_w, g_b, None, None, None


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
        self.train_precision = 3              # conv operands: 3 = 3xTF32 (fp32-grade), 4 = bf16, 1 = TF32

    def forward(self, img):
        """img [B,3,H,W] fp32 CUDA (NCHW) -> logits [B,1,H/16-1,W/16-1] (TRAIN:28-35)."""
        if not img.is_cuda:
            raise RuntimeError("PatchGAN (geomconsistentfr_b200) runs on CUDA only")
        if not self.training:
            raise NotImplementedError("the reference only ever runs the discriminator in train() mode (TRAIN:564-568)")
        prec = self.train_precision
        x = _SpaceToDepthPad.apply(img.float(), 3, True)                                  # [B,12,129,129]
        x = _ConvLReLU.apply(x, _w4x4_to_taps2x2(self.conv1.weight), self.conv1.bias, 12, 64, prec)
        for conv, bn, cin, cout in ((self.conv2, self.bn2, 64, 128), (self.conv3, self.bn3, 128, 256), (self.conv4, self.bn4, 256, 512)):
            x = _SpaceToDepthPad.apply(x, cin, False)
            meta = dict(cin=4 * cin, cout=cout, deconv=False, act=1, post_shift=0, bn=bn, taps=4, precision=prec)
            x = train_ops.ConvBNAct.apply(x, _w4x4_to_taps2x2(conv.weight), conv.bias, bn.weight, bn.bias, None, None, meta)
        return _Conv5.apply(x, self.conv5.weight, self.conv5.bias)
