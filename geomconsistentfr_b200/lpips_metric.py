"""LPIPS-Alex, spatial (the reference's third evaluation metric: PerceptualSimilarity/test_network.py:12-48 calls
`lpips.LPIPS(net='alex', spatial=True)` and averages the map over the face mask).

`LPIPSAlex` mirrors `lpips.LPIPS` (lpips/lpips.py:21-144) for net='alex', version '0.1', lpips=True: same sub-module /
parameter names, so `load_state_dict(torch.load('lpips/weights/v0.1/alex.pth'), strict=False)` installs the shipped linear
heads exactly as the reference does (lpips.py:109) and a torchvision AlexNet `features` state_dict fits `net.slice*`.
What runs where: the AlexNet trunk (5 convolutions, ReLU, 2 max-pools) is library code (cuDNN through torch) — its ImageNet
weights are not in this image, so it is only ever run with a seeded random trunk here; LPIPS's own arithmetic — channel
normalisation, squared difference, the learned 1x1 heads, bilinear upsampling to the input size, the masked mean — and its
backward are libgfr_b200 kernels (csrc/lpips.cu), differentiable w.r.t. both images ("LPIPS backward")."""
import torch
import torch.nn as nn

from . import _lib, ops
from .ops import _ptr, _stream

CHNS = (64, 192, 384, 256, 256)          # lpips.py:79


class _LayerDistance(torch.autograd.Function):
    """d [N,h,w] = sum_c w_c (normalize(f0) - normalize(f1))_c^2  (lpips.py:125-131 for one layer)."""

    @staticmethod
    def forward(ctx, f0, f1, w):
        f0, f1, w = f0.contiguous().float(), f1.contiguous().float(), w.reshape(-1).contiguous().float()
        N, C, h, wd = f0.shape
        out = torch.empty((N, h, wd), dtype=torch.float32, device=f0.device)
        _lib.check(_lib.load().gfr_lpips_layer_fwd(_ptr(f0), _ptr(f1), _ptr(w), _ptr(out), N, C, h * wd, _stream()), "gfr_lpips_layer_fwd")
        ops._count()
        ctx.save_for_backward(f0, f1, w)
        return out

    @staticmethod
    def backward(ctx, g):
        f0, f1, w = ctx.saved_tensors
        N, C, h, wd = f0.shape
        g0 = torch.empty_like(f0) if ctx.needs_input_grad[0] else None
        g1 = torch.empty_like(f1) if ctx.needs_input_grad[1] else None
        _lib.check(_lib.load().gfr_lpips_layer_bwd(_ptr(f0), _ptr(f1), _ptr(w), _ptr(g.contiguous()), _ptr(g0), _ptr(g1), N, C, h * wd,
                                                   _stream()), "gfr_lpips_layer_bwd")
        ops._count()
        return g0, g1, None


class _UpsampleAdd(torch.autograd.Function):
    """acc [N,H,W] + bilinear_upsample(m [N,h,w]) (nn.Upsample(size, 'bilinear', align_corners=False), lpips.py:16-18)."""

    @staticmethod
    def forward(ctx, acc, m):
        out = acc.clone()
        N, H, W = out.shape
        _, h, w = m.shape
        _lib.check(_lib.load().gfr_bilinear_up_add(_ptr(m.contiguous()), _ptr(out), N, h, w, H, W, _stream()), "gfr_bilinear_up_add")
        ops._count()
        ctx.shape = (N, h, w, H, W)
        return out

    @staticmethod
    def backward(ctx, g):
        N, h, w, H, W = ctx.shape
        gm = torch.zeros((N, h, w), dtype=torch.float32, device=g.device)
        _lib.check(_lib.load().gfr_bilinear_up_add_bwd(_ptr(g.contiguous()), _ptr(gm), N, h, w, H, W, _stream()), "gfr_bilinear_up_add_bwd")
        ops._count()
        return g, gm


class _ScalingLayer(nn.Module):                    # lpips.py:147-155
    def __init__(self):
        super().__init__()
        self.register_buffer("shift", torch.tensor([-.030, -.088, -.188])[None, :, None, None])
        self.register_buffer("scale", torch.tensor([.458, .448, .450])[None, :, None, None])

    def forward(self, x):
        return (x - self.shift) / self.scale


class _NetLinLayer(nn.Module):                     # lpips.py:158-169 (Dropout is the identity in eval mode)
    def __init__(self, chn_in):
        super().__init__()
        self.model = nn.Sequential(nn.Dropout(), nn.Conv2d(chn_in, 1, 1, stride=1, padding=0, bias=False))


class _AlexTrunk(nn.Module):
    """torchvision AlexNet `features` cut into the five slices of lpips/pretrained_networks.py:56-96 (same module indices)."""

    def __init__(self):
        super().__init__()
        self.slice1 = nn.Sequential()
        self.slice1.add_module("0", nn.Conv2d(3, 64, 11, stride=4, padding=2)); self.slice1.add_module("1", nn.ReLU(inplace=False))
        self.slice2 = nn.Sequential()
        self.slice2.add_module("2", nn.MaxPool2d(3, 2)); self.slice2.add_module("3", nn.Conv2d(64, 192, 5, padding=2)); self.slice2.add_module("4", nn.ReLU(inplace=False))
        self.slice3 = nn.Sequential()
        self.slice3.add_module("5", nn.MaxPool2d(3, 2)); self.slice3.add_module("6", nn.Conv2d(192, 384, 3, padding=1)); self.slice3.add_module("7", nn.ReLU(inplace=False))
        self.slice4 = nn.Sequential()
        self.slice4.add_module("8", nn.Conv2d(384, 256, 3, padding=1)); self.slice4.add_module("9", nn.ReLU(inplace=False))
        self.slice5 = nn.Sequential()
        self.slice5.add_module("10", nn.Conv2d(256, 256, 3, padding=1)); self.slice5.add_module("11", nn.ReLU(inplace=False))
        for p in self.parameters():                # requires_grad=False like lpips (pnet_tune=False)
            p.requires_grad_(False)

    def forward(self, x):
        outs = []
        for s in (self.slice1, self.slice2, self.slice3, self.slice4, self.slice5):
            x = s(x)
            outs.append(x)
        return outs


class LPIPSAlex(nn.Module):
    def __init__(self):
        super().__init__()
        self.scaling_layer = _ScalingLayer()
        self.net = _AlexTrunk()
        self.lin0, self.lin1, self.lin2, self.lin3, self.lin4 = (_NetLinLayer(c) for c in CHNS)
        self.lins = nn.ModuleList([self.lin0, self.lin1, self.lin2, self.lin3, self.lin4])       # lpips.py:93-99 registers them twice
        self.eval()

    def forward(self, in0, in1, normalize=False):
        """in0, in1 [N,3,H,W] in [-1,1] (or [0,1] with normalize=True) -> spatial distance map [N,1,H,W] (lpips.py:112-144,
        spatial=True, lpips=True)."""
        if not in0.is_cuda:
            raise RuntimeError("LPIPSAlex (geomconsistentfr_b200) runs on CUDA only")
        if normalize:
            in0, in1 = 2 * in0 - 1, 2 * in1 - 1
        N, _, H, W = in0.shape
        o0, o1 = self.net(self.scaling_layer(in0)), self.net(self.scaling_layer(in1))
        val = torch.zeros((N, H, W), dtype=torch.float32, device=in0.device)
        for k in range(5):
            d = _LayerDistance.apply(o0[k], o1[k], self.lins[k].model[1].weight)
            val = _UpsampleAdd.apply(val, d)
        return val.view(N, 1, H, W)


def masked_lpips(dist_map, mask):
    """test_network.py:41-45: sum(mask * map) / count(mask * map > 0) per image.  dist_map [N,1,H,W]; mask float [H,W] | [N,H,W]
    (the reference divides the mask image by 255) -> [N] f64."""
    m = dist_map.detach().reshape(dist_map.shape[0], dist_map.shape[2], dist_map.shape[3]).contiguous().float()
    N, H, W = m.shape
    mask = mask.to(m.device, torch.float32).contiguous()
    stride = 0 if mask.numel() == H * W else H * W
    sums = torch.empty((N, 2), dtype=torch.float64, device=m.device)
    _lib.check(_lib.load().gfr_lpips_masked_sums(_ptr(m), _ptr(mask), stride, _ptr(sums), N, H, W, _stream()), "gfr_lpips_masked_sums")
    ops._count()
    return sums[:, 0] / sums[:, 1]
