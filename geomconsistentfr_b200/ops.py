"""Operator layer: thin, checked wrappers that hand torch CUDA tensors to the C ABI (include/gfr_b200.h).

torch is used for device memory and the current stream only.  Every function raises if the input is not
a contiguous CUDA tensor of the expected dtype — there is no CPU path."""
import ctypes

import numpy as np
import torch

from . import _lib

NUM_SAMPLES = 160          # TRAIN:48
LIGHT_DISTANCE = 4013.0    # TRAIN:47
_MASK_DTYPES = {torch.uint8: 0, torch.bool: 0, torch.float32: 1, torch.float64: 2}


def reference_samples():
    """The reference's sample parameters, bit-for-bit: np.arange(0.025, 0.825, 0.005) (TRAIN:468)."""
    t = np.arange(0.025, 0.825, 0.005)
    assert t.shape[0] == NUM_SAMPLES
    return t


_launches = 0     # kernels launched through the C ABI (each entry point launches exactly one kernel)


def launch_count():
    return _launches


def _count(n=1):
    global _launches
    _launches += n


_param_generation = 0    # bumped by everything that writes parameters / BatchNorm buffers through raw pointers (fused Adam, the
                         # batch-statistics kernels, graph replays of a training step): torch's `_version` counters do not see
                         # those writes, so caches of derived weights (BN-folded, packed) key on this as well


def param_generation():
    return _param_generation


def bump_param_generation():
    global _param_generation
    _param_generation += 1


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need(t, dtype, name):
    if not (torch.is_tensor(t) and t.is_cuda):
        raise RuntimeError("%s must be a CUDA tensor (libgfr_b200 has no CPU path)" % name)
    if t.dtype != dtype:
        raise RuntimeError("%s must be %s, got %s" % (name, dtype, t.dtype))
    return t if t.is_contiguous() else t.contiguous()


MASK_EXTRA_WORDS = 4      # GFR_MASK_EXTRA_WORDS: the bounding box of the mask rides behind its bitmap


def mask_pack(mask):
    """mask [n,H,W] (u8/bool/f32/f64, CUDA) -> bits [n, H*W/32 + 4] int32 (bit set iff mask != 0, TRAIN:510; + bbox)."""
    if not (torch.is_tensor(mask) and mask.is_cuda):
        raise RuntimeError("mask must be a CUDA tensor")
    if mask.dtype not in _MASK_DTYPES:
        raise RuntimeError("unsupported mask dtype %s" % mask.dtype)
    mask = mask.contiguous()
    n, H, W = mask.shape
    bits = torch.empty((n, H * W // 32 + MASK_EXTRA_WORDS), dtype=torch.int32, device=mask.device)
    rc = _lib.load().gfr_mask_pack(_ptr(mask), _MASK_DTYPES[mask.dtype], n, H, W, _ptr(bits), _stream())
    _lib.check(rc, "gfr_mask_pack"); _count()
    return bits


def _bonus_rect(rect):
    """None -> NULL (the image rectangle, TEST1:495); else 4 host floats {x_min, x_max, y_min, y_max} (TEST_LT:503)."""
    if rect is None:
        return None, None
    r = np.ascontiguousarray(rect, dtype=np.float32).reshape(4)
    return r, r.ctypes.data_as(ctypes.c_void_p)


def shadow_march_fwd(depth, mask_bits, light_pt, samples=None, inside_bonus=0.0, want_argmin=False,
                     want_shadow=False, variant=0, bonus_rect=None):
    """depth [F,1,H,W] f32; mask_bits [1|F, H*W/32] i32; light_pt [B,3] f32 with B = F * L (L lights per face, pair b
    uses face b // L) -> d_min [B,H,W] (+ argmin u8, + shadow).  TRAIN:374-517."""
    depth = _need(depth, torch.float32, "depth")
    light_pt = _need(light_pt, torch.float32, "light_pt")
    mask_bits = _need(mask_bits, torch.int32, "mask_bits")
    F, _, H, W = depth.shape
    B = light_pt.shape[0]
    if light_pt.dim() != 2 or light_pt.shape[1] != 3 or B % F:
        raise RuntimeError("light_pt must be [F*L,3]")
    if mask_bits.shape[0] not in (1, F) or mask_bits.shape[1] != H * W // 32 + MASK_EXTRA_WORDS:
        raise RuntimeError("mask_bits must be [1|F, H*W/32 + 4] (from mask_pack)")
    t = reference_samples() if samples is None else np.ascontiguousarray(samples, dtype=np.float64)
    dmin = torch.empty((B, H, W), dtype=torch.float32, device=depth.device)
    arg = torch.empty((B, H, W), dtype=torch.uint8, device=depth.device) if want_argmin else None
    shadow = torch.empty((B, H, W), dtype=torch.float32, device=depth.device) if want_shadow else None
    stride = 0 if mask_bits.shape[0] == 1 else H * W // 32 + MASK_EXTRA_WORDS
    scratch = torch.empty(F * H * W + F, dtype=torch.float64, device=depth.device) if variant in (0, 2) else None    # fp64 depth + the per-face depth range
    _keep, rect = _bonus_rect(bonus_rect)
    rc = _lib.load().gfr_shadow_march_fwd(
        _ptr(depth), _ptr(mask_bits), stride, _ptr(light_pt), t.ctypes.data_as(ctypes.c_void_p), int(t.shape[0]),
        float(inside_bonus), rect, _ptr(dmin), _ptr(arg), _ptr(shadow), _ptr(scratch), B, H, W, B // F, int(variant), _stream())
    _lib.check(rc, "gfr_shadow_march_fwd"); _count(2 if variant in (0, 2) else 1)
    return dmin, arg, shadow


def march_config(warp_shape=-1, ilp=0, block_order=-1, early_cut=-1):
    """A/B switches of the default march kernel (process-wide): warp_shape 0 = 32x1, 1 = 8x4 pixels per warp; ilp 1..4 =
    samples one by one / in groups; block_order 0 = tile-major, 1 = pairs interleaved + far-from-light tiles first;
    early_cut 0 / 1 = walk every in-range sample / stop a ray once no later sample can beat its minimum (sample groups only);
    -1 / 0 = defaults.  Bit-identical results in every setting."""
    _lib.check(_lib.load().gfr_march_config(int(warp_shape), int(ilp), int(block_order), int(early_cut)), "gfr_march_config")


def shade_render_fwd(albedo, depth, d_min, light_pt, ambient, fx=1570.0, fy=1570.0, cx=None, cy=None,
                     depth_offset=1610.0, intensity=0.5, want=("shadow", "full", "final", "rendered", "normals")):
    """TRAIN:353-369, 517-522.  albedo/depth/ambient hold F faces, d_min/light_pt B = F * L (face, light) pairs.
    Returns dict of the requested outputs ([B,...])."""
    depth = _need(depth, torch.float32, "depth")
    d_min = _need(d_min, torch.float32, "d_min")
    light_pt = _need(light_pt, torch.float32, "light_pt")
    ambient = _need(ambient.reshape(-1), torch.float32, "ambient")
    F, _, H, W = depth.shape
    B = d_min.shape[0]
    if B % F or light_pt.shape != (B, 3) or ambient.shape[0] != F:
        raise RuntimeError("shade_render_fwd: d_min/light_pt must hold F*L pairs, ambient F faces")
    if "rendered" in want:
        albedo = _need(albedo, torch.float32, "albedo")
    intr = np.array([fx, fy, W / 2.0 if cx is None else cx, H / 2.0 if cy is None else cy, depth_offset, intensity],
                    dtype=np.float32)
    dev = depth.device
    out = {}
    for k, shape in (("shadow", (B, H, W)), ("full", (B, H, W)), ("final", (B, H, W)), ("rendered", (B, 3, H, W)),
                     ("normals", (B, 3, H, W))):
        out[k] = torch.empty(shape, dtype=torch.float32, device=dev) if k in want else None
    rc = _lib.load().gfr_shade_render_fwd(
        _ptr(albedo) if "rendered" in want else None, _ptr(depth), _ptr(d_min), _ptr(light_pt), _ptr(ambient),
        intr.ctypes.data_as(ctypes.c_void_p), _ptr(out["shadow"]), _ptr(out["full"]), _ptr(out["final"]),
        _ptr(out["rendered"]), _ptr(out["normals"]), B, H, W, B // F, _stream())
    _lib.check(rc, "gfr_shade_render_fwd"); _count()
    return out


def march_shade_fwd(albedo, depth, mask_bits, light_pt, ambient, inside_bonus=0.0, fx=1570.0, fy=1570.0, cx=None, cy=None,
                    depth_offset=1610.0, intensity=0.5, want=("shadow", "full", "final", "rendered", "normals"),
                    samples=None, bonus_rect=None):
    """shadow_march_fwd + shade_render_fwd in one launch (d_min stays on the SM).  Same arguments / outputs as the pair;
    returns the dict of the requested outputs ([B,...], B = F * lights per face)."""
    depth = _need(depth, torch.float32, "depth")
    light_pt = _need(light_pt, torch.float32, "light_pt")
    mask_bits = _need(mask_bits, torch.int32, "mask_bits")
    ambient = _need(ambient.reshape(-1), torch.float32, "ambient")
    F, _, H, W = depth.shape
    B = light_pt.shape[0]
    if light_pt.dim() != 2 or light_pt.shape[1] != 3 or B % F or ambient.shape[0] != F:
        raise RuntimeError("march_shade_fwd: light_pt must be [F*L,3], ambient [F]")
    if mask_bits.shape[0] not in (1, F) or mask_bits.shape[1] != H * W // 32 + MASK_EXTRA_WORDS:
        raise RuntimeError("mask_bits must be [1|F, H*W/32 + 4] (from mask_pack)")
    if "rendered" in want:
        albedo = _need(albedo, torch.float32, "albedo")
    t = reference_samples() if samples is None else np.ascontiguousarray(samples, dtype=np.float64)
    _keep, rect = _bonus_rect(bonus_rect)
    intr = np.array([fx, fy, W / 2.0 if cx is None else cx, H / 2.0 if cy is None else cy, depth_offset, intensity],
                    dtype=np.float32)
    dev = depth.device
    out = {}
    for k, shape in (("shadow", (B, H, W)), ("full", (B, H, W)), ("final", (B, H, W)), ("rendered", (B, 3, H, W)),
                     ("normals", (B, 3, H, W)), ("d_min", (B, H, W))):
        out[k] = torch.empty(shape, dtype=torch.float32, device=dev) if k in want else None
    stride = 0 if mask_bits.shape[0] == 1 else H * W // 32 + MASK_EXTRA_WORDS
    scratch = torch.empty(F * H * W + F, dtype=torch.float64, device=dev)       # fp64 depth + the per-face depth range
    rc = _lib.load().gfr_march_shade_fwd(
        _ptr(albedo) if "rendered" in want else None, _ptr(depth), _ptr(mask_bits), stride, _ptr(light_pt), _ptr(ambient),
        t.ctypes.data_as(ctypes.c_void_p), int(t.shape[0]), float(inside_bonus), rect, intr.ctypes.data_as(ctypes.c_void_p),
        _ptr(scratch), _ptr(out["d_min"]), None, _ptr(out["shadow"]), _ptr(out["full"]), _ptr(out["final"]),
        _ptr(out["rendered"]), _ptr(out["normals"]), B, H, W, B // F, _stream())
    _lib.check(rc, "gfr_march_shade_fwd"); _count(2)
    return out


# ---------------------------------------------------------------------------------------------------
# CNN building blocks (fp32 direct path)
# ---------------------------------------------------------------------------------------------------
_ACT = {None: 0, "none": 0, "lrelu": 1, "sigmoid": 2}


def conv2d_fwd(x, w, bias, res=None, post=None, post_shift=0, act="lrelu", ups_in=False, out_scale=1.0):
    """out = out_scale * (act(conv(x') + bias + res) + up(post)).  x: logical NCHW view (any strides);
    w [Cout,Cin,K,K] (BN folded); see include/gfr_b200.h gfr_conv2d_fwd."""
    if not (x.is_cuda and x.dtype == torch.float32):
        raise RuntimeError("conv2d_fwd: x must be a CUDA fp32 tensor")
    w = _need(w, torch.float32, "w")
    bias = _need(bias, torch.float32, "bias")
    N, Cin, Hin, Win = x.shape
    H, W = (Hin * 2, Win * 2) if ups_in else (Hin, Win)
    Cout, Cin_w, K, _ = w.shape
    if Cin_w != Cin:
        raise RuntimeError("conv2d_fwd: Cin mismatch")
    if res is not None:
        res = _need(res, torch.float32, "res")
        assert res.shape == (N, Cout, H, W)
    if post is not None:
        post = _need(post, torch.float32, "post")
        assert post.shape == (N, Cout, H >> post_shift, W >> post_shift)
    out = torch.empty((N, Cout, H, W), dtype=torch.float32, device=x.device)
    strides = (ctypes.c_longlong * 4)(*x.stride())
    rc = _lib.load().gfr_conv2d_fwd(_ptr(x), ctypes.cast(strides, ctypes.c_void_p), _ptr(w), _ptr(bias), _ptr(res),
                                    _ptr(post), _ptr(out), N, Cin, Cout, H, W, K, int(bool(ups_in)), int(post_shift),
                                    _ACT[act], float(out_scale), _stream())
    _lib.check(rc, "gfr_conv2d_fwd"); _count()
    return out


def maxpool2_fwd(x):
    x = _need(x, torch.float32, "x")
    N, C, H, W = x.shape
    out = torch.empty((N, C, H // 2, W // 2), dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().gfr_maxpool2_fwd(_ptr(x), _ptr(out), N * C, H // 2, W // 2, _stream()), "gfr_maxpool2_fwd")
    _count()
    return out


def upsample2_fwd(x, add=None):
    x = _need(x, torch.float32, "x")
    N, C, H, W = x.shape
    if add is not None:
        add = _need(add, torch.float32, "add")
    out = torch.empty((N, C, 2 * H, 2 * W), dtype=torch.float32, device=x.device)
    _lib.check(_lib.load().gfr_upsample2_fwd(_ptr(x), _ptr(add), _ptr(out), N * C, 2 * H, 2 * W, _stream()),
               "gfr_upsample2_fwd")
    _count()
    return out


def light_head_fwd(feat, c_first, w1, b1, w2, b2):
    """feat [N,C,h,w] contiguous; channels [c_first, c_first+27) -> [N,4] (TRAIN:225-232)."""
    feat = _need(feat, torch.float32, "feat")
    N, C, h, w = feat.shape
    out = torch.empty((N, 4), dtype=torch.float32, device=feat.device)
    rc = _lib.load().gfr_light_head_fwd(_ptr(feat), C * h * w, int(c_first), h * w, _ptr(_need(w1, torch.float32, "w1")),
                                        _ptr(_need(b1, torch.float32, "b1")), _ptr(_need(w2, torch.float32, "w2")),
                                        _ptr(_need(b2, torch.float32, "b2")), _ptr(out), N, _stream())
    _lib.check(rc, "gfr_light_head_fwd"); _count()
    return out


# ---------------------------------------------------------------------------------------------------
# tensor-core CNN path: C4 activation layout [N][ceil(C/4)][H][W][4]
# ---------------------------------------------------------------------------------------------------
class C4:
    """An activation tensor in the C4 layout: data [N, ceil(C/4), H, W, 4] fp32 CUDA + the logical channel count."""
    __slots__ = ("data", "C")

    def __init__(self, data, C):
        self.data, self.C = data, C

    @property
    def shape(self):
        N, _, H, W, _ = self.data.shape
        return (N, self.C, H, W)


def _c4_empty(N, C, H, W, device):
    return torch.empty((N, (C + 3) // 4, H, W, 4), dtype=torch.float32, device=device)


def nchw_to_c4(x):
    x = _need(x, torch.float32, "x")
    N, C, H, W = x.shape
    out = _c4_empty(N, C, H, W, x.device)
    _lib.check(_lib.load().gfr_nchw_to_c4(_ptr(x), _ptr(out), N, C, H, W, _stream()), "gfr_nchw_to_c4"); _count()
    return C4(out, C)


def c4_to_nchw(x):
    N, C, H, W = x.shape
    out = torch.empty((N, C, H, W), dtype=torch.float32, device=x.data.device)
    _lib.check(_lib.load().gfr_c4_to_nchw(_ptr(x.data), _ptr(out), N, C, H, W, _stream()), "gfr_c4_to_nchw"); _count()
    return out


def conv_tc_pack_weights(w, NT):
    """w [Cout,Cin,3,3] fp32 (any device) -> packed CUDA/CPU tensor on w's device (host-side packing, model-load time)."""
    wc = w.detach().to("cpu", torch.float32).contiguous()
    Cout, Cin, K, _ = wc.shape
    assert K == 3
    n = _lib.load().gfr_conv_tc_pack_size(Cin, Cout, NT)
    if n < 0:
        raise RuntimeError("gfr_conv_tc_pack_size: bad arguments")
    packed = torch.empty(n, dtype=torch.float32)
    rc = _lib.load().gfr_conv_tc_pack_weights(ctypes.c_void_p(wc.data_ptr()), Cin, Cout, NT, ctypes.c_void_p(packed.data_ptr()))
    _lib.check(rc, "gfr_conv_tc_pack_weights")
    return packed.to(w.device)


X_SCALE_F16 = 16.0     # activations are multiplied by 2^4 before the fp16 pair split: |x| < 4095 stays finite (RelightNet: |x| < 25),
                       # and the fp16-subnormal floor of the correction term is 3e-8 / 16 = 2e-9 absolute


def conv_tc_pack_weights_f16(w):
    """w [Cout,Cin,3,3] -> (packed, w_scale, NT) for precision=2 (fp16 pair split); w_scale = the largest power of two
    with max|w| * w_scale <= 2^14."""
    wc = w.detach().to("cpu", torch.float32).contiguous()
    Cout, Cin, K, _ = wc.shape
    assert K == 3
    NT = 16 if Cout <= 16 else 32
    mx = float(wc.abs().max())
    w_scale = 2.0 ** min(12, int(np.floor(np.log2(16384.0 / max(mx, 1e-30)))))
    n = _lib.load().gfr_conv_tc_pack_size_f16(Cin, Cout, NT)
    packed = torch.empty(n, dtype=torch.float32)
    rc = _lib.load().gfr_conv_tc_pack_weights_f16(ctypes.c_void_p(wc.data_ptr()), Cin, Cout, NT, ctypes.c_float(w_scale),
                                                  ctypes.c_void_p(packed.data_ptr()))
    _lib.check(rc, "gfr_conv_tc_pack_weights_f16")
    return packed.to(w.device), w_scale, NT


def conv3x3_tc_fwd(x, w_packed, bias, Cout, NT, res=None, post=None, post_shift=0, act="lrelu", out_scale=1.0,
                   precision=3, cin=None, w_scale=1.0):
    """x: C4; w_packed from conv_tc_pack_weights(w, NT); out = out_scale*(act(conv(x)+bias+res) + up(post)) as C4.
    cin < x.C reads only the leading cin channels of x (TRAIN:225 feature split), in place."""
    N, Cin, H, W = x.shape
    groups = x.data.shape[1]
    if cin is not None:
        assert cin <= Cin and cin % 4 == 0
        Cin = cin
    w_packed = _need(w_packed, torch.float32, "w_packed")
    bias = _need(bias, torch.float32, "bias")
    if not x.data.is_cuda:
        raise RuntimeError("conv3x3_tc_fwd: x must be on CUDA")
    out = _c4_empty(N, Cout, H, W, x.data.device)
    if res is not None:
        assert res.shape == (N, Cout, H, W)
    if post is not None:
        assert post.shape == (N, Cout, H >> post_shift, W >> post_shift)
    rc = _lib.load().gfr_conv3x3_tc_fwd(_ptr(x.data), _ptr(w_packed), _ptr(bias), _ptr(res.data if res is not None else None),
                                        _ptr(post.data if post is not None else None), _ptr(out), N, Cin, groups, Cout, H, W, NT,
                                        int(post_shift), _ACT[act], float(out_scale), int(precision), 1, X_SCALE_F16, float(w_scale),
                                        _stream())
    _lib.check(rc, "gfr_conv3x3_tc_fwd"); _count()
    return C4(out, Cout)


def pack_head_weights(w2, b2, w3, b3, wo, bo, device):
    """The decoder tail's six (BN-folded) tensors as the 600-float device operand of gfr_conv3x3_tc_head_fwd:
    [w2 16x16 | b2 16 | w3 16x16 | b3 16 | wo 3x16 | bo 4], rows [co][ci]."""
    n_out = wo.shape[0]
    buf = torch.zeros(600, dtype=torch.float32)
    buf[0:256] = w2.reshape(16, 16).reshape(-1)
    buf[256:272] = b2
    buf[272:528] = w3.reshape(16, 16).reshape(-1)
    buf[528:544] = b3
    buf[544:544 + 16 * n_out] = wo.reshape(n_out, 16).reshape(-1)
    buf[592:592 + n_out] = bo
    return buf.to(device)


def conv3x3_tc_head_fwd(x, w_packed, bias, head, n_out, act=None, out_scale=1.0, precision=2, w_scale=1.0):
    """c2_1 (3x3, 16 -> 16, BN folded, LeakyReLU) with the 1x1 tail c2_2 -> c2_3 -> c2_o fused into its epilogue
    (TRAIN:284-290 / 344-350).  x: C4 [N,16,H,W]; head from pack_head_weights -> NCHW [N,n_out,H,W]."""
    N, Cin, H, W = x.shape
    if not x.data.is_cuda:
        raise RuntimeError("conv3x3_tc_head_fwd: x must be on CUDA")
    w_packed, bias, head = _need(w_packed, torch.float32, "w_packed"), _need(bias, torch.float32, "bias"), _need(head, torch.float32, "head")
    if head.numel() != 600:
        raise RuntimeError("head must come from pack_head_weights (600 floats)")
    out = torch.empty((N, n_out, H, W), dtype=torch.float32, device=x.data.device)
    rc = _lib.load().gfr_conv3x3_tc_head_fwd(_ptr(x.data), _ptr(w_packed), _ptr(bias), _ptr(head), _ptr(out), N, Cin,
                                             x.data.shape[1], H, W, int(n_out), _ACT[act], float(out_scale), int(precision), 1,
                                             X_SCALE_F16, float(w_scale), _stream())
    _lib.check(rc, "gfr_conv3x3_tc_head_fwd"); _count()
    return out


def maxpool2_c4_fwd(x):
    N, C, H, W = x.shape
    out = _c4_empty(N, C, H // 2, W // 2, x.data.device)
    _lib.check(_lib.load().gfr_maxpool2_c4_fwd(_ptr(x.data), _ptr(out), N * ((C + 3) // 4), H // 2, W // 2, _stream()),
               "gfr_maxpool2_c4_fwd"); _count()
    return C4(out, C)


def upsample2_c4_fwd(x, add=None):
    N, C, H, W = x.shape
    out = _c4_empty(N, C, 2 * H, 2 * W, x.data.device)
    _lib.check(_lib.load().gfr_upsample2_c4_fwd(_ptr(x.data), _ptr(add.data if add is not None else None), _ptr(out),
                                                N * ((C + 3) // 4), 2 * H, 2 * W, _stream()), "gfr_upsample2_c4_fwd"); _count()
    return C4(out, C)


def stem_conv_fwd(img, w_host, b_host, pool=True):
    """img [N,H,W,3] fp32 CUDA; w_host [16,3,5,5], b_host [16] CPU tensors (BN folded) -> (C4 [N,16,H,W], pooled C4)."""
    img = _need(img, torch.float32, "img")
    N, H, W, _ = img.shape
    out = _c4_empty(N, 16, H, W, img.device)
    pooled = _c4_empty(N, 16, H // 2, W // 2, img.device) if pool else None
    rc = _lib.load().gfr_stem_conv_fwd(_ptr(img), ctypes.c_void_p(w_host.data_ptr()), ctypes.c_void_p(b_host.data_ptr()),
                                       _ptr(out), _ptr(pooled), N, H, W, _stream())
    _lib.check(rc, "gfr_stem_conv_fwd"); _count()
    return C4(out, 16), (C4(pooled, 16) if pool else None)


def head_1x1_fwd(x, w2, b2, w3, b3, wo, bo, act=None, out_scale=1.0):
    """x C4 [N,16,H,W]; weights are CPU tensors (BN folded) -> NCHW [N,n_out,H,W]."""
    N, C, H, W = x.shape
    assert C == 16
    n_out = wo.shape[0]
    out = torch.empty((N, n_out, H, W), dtype=torch.float32, device=x.data.device)
    hp = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = _lib.load().gfr_head_1x1_fwd(_ptr(x.data), hp(w2), hp(b2), hp(w3), hp(b3), hp(wo), hp(bo), _ptr(out), N, H, W, n_out,
                                      _ACT[act], float(out_scale), _stream())
    _lib.check(rc, "gfr_head_1x1_fwd"); _count()
    return out


def light_head_c4_fwd(feat, c_first, w1, b1, w2, b2):
    """feat C4 [N,C,h,w]; channels [c_first, c_first+27) -> [N,4] (TRAIN:225-232)."""
    N, C, h, w = feat.shape
    out = torch.empty((N, 4), dtype=torch.float32, device=feat.data.device)
    rc = _lib.load().gfr_light_head_c4_fwd(_ptr(feat.data), C, int(c_first), h * w, _ptr(_need(w1, torch.float32, "w1")),
                                           _ptr(_need(b1, torch.float32, "b1")), _ptr(_need(w2, torch.float32, "w2")),
                                           _ptr(_need(b2, torch.float32, "b2")), _ptr(out), N, _stream())
    _lib.check(rc, "gfr_light_head_c4_fwd"); _count()
    return out


# ---------------------------------------------------------------------------------------------------
# eval-mode tensor-core CNN on PRE-SPLIT fp16-pair activations: the P16 layout [N][ceil(C/8)][hi|lo][H][W][8] fp16
# (csrc/p16.cuh, csrc/conv_p16.cu)
# ---------------------------------------------------------------------------------------------------
class P16:
    """An activation tensor in the P16 layout: data [N, groups, 2, H, W, 8] fp16 CUDA + the logical channel count."""
    __slots__ = ("data", "C")

    def __init__(self, data, C):
        self.data, self.C = data, C

    @property
    def shape(self):
        N, _, _, H, W, _ = self.data.shape
        return (N, self.C, H, W)

    @property
    def groups(self):
        return self.data.shape[1]


def _p16_empty(N, C, H, W, device):
    return torch.empty((N, (C + 7) // 8, 2, H, W, 8), dtype=torch.float16, device=device)


def nchw_to_p16(x):
    x = _need(x, torch.float32, "x")
    N, C, H, W = x.shape
    out = _p16_empty(N, C, H, W, x.device)
    _lib.check(_lib.load().gfr_nchw_to_p16(_ptr(x), _ptr(out), N, C, H, W, _stream()), "gfr_nchw_to_p16"); _count()
    return P16(out, C)


def p16_to_nchw(x, C=None, c_first=0):
    """P16 -> NCHW fp32; `c_first` (a multiple of 8) / `C` select a channel range of a wider tensor."""
    N, Cx, H, W = x.shape
    C = Cx - c_first if C is None else C
    assert c_first % 8 == 0
    out = torch.empty((N, C, H, W), dtype=torch.float32, device=x.data.device)
    base = x.data[:, c_first // 8:]
    rc = _lib.load().gfr_p16_to_nchw(ctypes.c_void_p(base.data_ptr()), _ptr(out), N, C, x.groups, H, W, _stream())
    _lib.check(rc, "gfr_p16_to_nchw"); _count()
    return out


def conv_p16_config(Cin, Cout, N, H, W):
    """(NT, MH, KS) for a layer: 16-wide N tiles for <= 16 output channels, 16x16-pixel tiles when the image gives every
    SM several of them, 32 input channels per pipeline step when there are that many."""
    NT = 16 if Cout <= 16 else 32
    MH = 2 if (N * ((H + 15) // 16) * ((W + 15) // 16)) * ((Cout + NT - 1) // NT) >= 2 * 148 else 1
    KS = 4 if Cin >= 32 else 2
    return NT, MH, KS


def conv_p16_pack_weights(w, NT, KS):
    """w [Cout,Cin,3,3] fp32 -> (packed fp16 tensor on w's device, w_scale): w * w_scale = w1 + w2 as an fp16 pair, w_scale =
    the largest power of two with max|w| * w_scale <= 2^14 (host-side packing, model-load time)."""
    wc = w.detach().to("cpu", torch.float32).contiguous()
    Cout, Cin, K, _ = wc.shape
    assert K == 3
    mx = float(wc.abs().max())
    w_scale = 2.0 ** min(12, int(np.floor(np.log2(16384.0 / max(mx, 1e-30)))))
    n = _lib.load().gfr_conv_p16_pack_size(Cin, Cout, NT, KS)
    if n < 0:
        raise RuntimeError("gfr_conv_p16_pack_size: bad arguments")
    packed = torch.empty(n, dtype=torch.float16)
    rc = _lib.load().gfr_conv_p16_pack_weights(ctypes.c_void_p(wc.data_ptr()), Cin, Cout, NT, KS, ctypes.c_float(w_scale),
                                               ctypes.c_void_p(packed.data_ptr()))
    _lib.check(rc, "gfr_conv_p16_pack_weights")
    return packed.to(w.device), w_scale


def conv_p16_pack_weights_taps(w, NT, KS):
    """w [Cout,Cin,taps] (taps = 5: the vertical taps of the unrolled stem) -> (packed fp16 tensor, w_scale)."""
    wc = w.detach().to("cpu", torch.float32).contiguous()
    Cout, Cin, taps = wc.shape
    mx = float(wc.abs().max())
    w_scale = 2.0 ** min(12, int(np.floor(np.log2(16384.0 / max(mx, 1e-30)))))
    n = _lib.load().gfr_conv_p16_pack_size_taps(Cin, Cout, NT, KS, taps)
    if n < 0:
        raise RuntimeError("gfr_conv_p16_pack_size_taps: bad arguments")
    packed = torch.empty(n, dtype=torch.float16)
    rc = _lib.load().gfr_conv_p16_pack_weights_taps(ctypes.c_void_p(wc.data_ptr()), Cin, Cout, NT, KS, taps, ctypes.c_float(w_scale),
                                                    ctypes.c_void_p(packed.data_ptr()))
    _lib.check(rc, "gfr_conv_p16_pack_weights_taps")
    return packed.to(w.device), w_scale


def stem_unroll_p16(img):
    """img [N,H,W,3] fp32 CUDA -> P16 [N,16,H,W]: channel kx*3 + c = img[..., x + kx - 2, c] (the stem's horizontal taps as channels)."""
    img = _need(img, torch.float32, "img")
    N, H, W, _ = img.shape
    out = _p16_empty(N, 16, H, W, img.device)
    _lib.check(_lib.load().gfr_stem_unroll_p16(_ptr(img), _ptr(out), N, H, W, _stream()), "gfr_stem_unroll_p16"); _count()
    return P16(out, 16)


def conv3x3_p16_fwd(x, w_packed, bias, Cout, cfg, w_scale, res=None, res_c=0, post=None, post_shift=0, act="lrelu",
                    act_channels=0, out_scale=1.0, cin=None, flags=None, pool=False, geometry=0, cin_first=0):
    """x: P16; w_packed / w_scale from conv_p16_pack_weights(w, NT, KS); cfg = (NT, MH, KS).
    out = out_scale * (act(conv(x[:, :cin]) + bias + res[:, res_c:res_c+Cout]) + up(post)) as P16 (the activation only on
    output channels < act_channels when given).  `flags`: int32[1] CUDA tensor that collects range overflows."""
    N, Cin, H, W = x.shape
    if cin is not None:
        assert cin_first + cin <= x.groups * 8
        Cin = cin
    assert cin_first % 8 == 0 and (cin_first == 0 or cin is not None)
    NT, MH, KS = cfg
    if not x.data.is_cuda:
        raise RuntimeError("conv3x3_p16_fwd: x must be on CUDA")
    # the input channels [cin_first, cin_first + Cin) of a wider tensor: the chunk offset goes into the base pointer, the
    # stride between images stays that of the whole tensor (in_groups)
    x_ptr = ctypes.c_void_p(x.data.data_ptr() + (cin_first // 8) * 2 * H * W * 8 * 2)
    w_packed = _need(w_packed, torch.float16, "w_packed")
    bias = _need(bias, torch.float32, "bias")
    out = _p16_empty(N, Cout, H, W, x.data.device)
    if res is not None:
        assert res.shape[0] == N and res.shape[2:] == (H, W) and res_c % 8 == 0 and res_c + Cout <= res.groups * 8
    if post is not None:
        assert post.shape == (N, Cout, H >> post_shift, W >> post_shift)
    pooled = _p16_empty(N, Cout, H // 2, W // 2, x.data.device) if pool else None
    rc = _lib.load().gfr_conv_p16_fwd_ex(
        x_ptr, _ptr(w_packed), _ptr(bias), _ptr(res.data if res is not None else None), res_c // 8,
        res.groups if res is not None else 0, _ptr(post.data if post is not None else None), post.groups if post is not None else 0,
        _ptr(out), 0, _ptr(pooled), _ptr(flags), N, Cin, x.groups, Cout, H, W, NT, MH, KS, int(geometry), int(post_shift), _ACT[act],
        int(act_channels), float(out_scale), float(w_scale), 1, _stream())
    _lib.check(rc, "gfr_conv_p16_fwd_ex"); _count()
    if pool:
        return P16(out, Cout), P16(pooled, Cout)
    return P16(out, Cout)


def maxpool2_p16_fwd(x):
    N, C, H, W = x.shape
    out = _p16_empty(N, C, H // 2, W // 2, x.data.device)
    _lib.check(_lib.load().gfr_maxpool2_p16_fwd(_ptr(x.data), _ptr(out), N * x.groups, H // 2, W // 2, _stream()),
               "gfr_maxpool2_p16_fwd"); _count()
    return P16(out, C)


def upsample2_p16_fwd(x):
    """nearest x2: the hi and lo planes are upsampled as independent planes of 16-byte units (the C4 kernel does that)."""
    N, C, H, W = x.shape
    out = _p16_empty(N, C, 2 * H, 2 * W, x.data.device)
    _lib.check(_lib.load().gfr_upsample2_c4_fwd(_ptr(x.data), None, _ptr(out), N * x.groups * 2, 2 * H, 2 * W, _stream()),
               "gfr_upsample2_c4_fwd"); _count()
    return P16(out, C)


def stem_conv_p16_fwd(img, w_host, b_host, pool=True):
    """img [N,H,W,3] fp32 CUDA; w_host [16,3,5,5], b_host [16] CPU tensors (BN folded) -> (P16 [N,16,H,W], pooled P16)."""
    img = _need(img, torch.float32, "img")
    N, H, W, _ = img.shape
    out = _p16_empty(N, 16, H, W, img.device)
    pooled = _p16_empty(N, 16, H // 2, W // 2, img.device) if pool else None
    rc = _lib.load().gfr_stem_conv_p16_fwd(_ptr(img), ctypes.c_void_p(w_host.data_ptr()), ctypes.c_void_p(b_host.data_ptr()),
                                           _ptr(out), _ptr(pooled), N, H, W, _stream())
    _lib.check(rc, "gfr_stem_conv_p16_fwd"); _count()
    return P16(out, 16), (P16(pooled, 16) if pool else None)


def head_1x1_p16_fwd(x, w2, b2, w3, b3, wo, bo, act=None, out_scale=1.0):
    """x P16 [N,16,H,W]; weights are CPU tensors (BN folded) -> NCHW [N,n_out,H,W]."""
    N, C, H, W = x.shape
    assert C == 16 and x.groups == 2
    n_out = wo.shape[0]
    out = torch.empty((N, n_out, H, W), dtype=torch.float32, device=x.data.device)
    hp = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = _lib.load().gfr_head_1x1_p16_fwd(_ptr(x.data), hp(w2), hp(b2), hp(w3), hp(b3), hp(wo), hp(bo), _ptr(out), N, H, W, n_out,
                                          _ACT[act], float(out_scale), _stream())
    _lib.check(rc, "gfr_head_1x1_p16_fwd"); _count()
    return out


def conv_p16_config(grid_ctas_per_sm=0):
    """Persistent grid of the P16 convolutions: 0 default, 1 = one CTA per SM (throughput), 2 = two where they fit (latency)."""
    _lib.check(_lib.load().gfr_conv_p16_config(int(grid_ctas_per_sm)), "gfr_conv_p16_config")


def conv3x3_p16_head_fwd(x, w_packed, bias, MH, w_scale, w2, b2, w3, b3, wo, bo, act=None, out_scale=1.0, cin=None):
    """conv3x3_p16_fwd (Cin -> 16, LeakyReLU, cfg (16, MH, 2)) + head_1x1_p16_fwd in ONE launch (the 1x1 tail runs in the conv's
    epilogue) -> NCHW [N,n_out,H,W]."""
    N, Cin, H, W = x.shape
    if cin is not None:
        assert cin <= Cin
        Cin = cin
    if not x.data.is_cuda:
        raise RuntimeError("conv3x3_p16_head_fwd: x must be on CUDA")
    w_packed = _need(w_packed, torch.float16, "w_packed")
    bias = _need(bias, torch.float32, "bias")
    n_out = wo.shape[0]
    out = torch.empty((N, n_out, H, W), dtype=torch.float32, device=x.data.device)
    hp = lambda t: ctypes.c_void_p(t.data_ptr())
    rc = _lib.load().gfr_conv3x3_p16_head_fwd(_ptr(x.data), _ptr(w_packed), _ptr(bias), N, Cin, x.groups, H, W, int(MH), float(w_scale), 1,
                                              hp(w2), hp(b2), hp(w3), hp(b3), hp(wo), hp(bo), _ptr(out), n_out, _ACT[act],
                                              float(out_scale), _stream())
    _lib.check(rc, "gfr_conv3x3_p16_head_fwd"); _count()
    return out


def light_head_p16_fwd(feat, c_first, w1, b1, w2, b2):
    """feat P16 [N,C,h,w]; channels [c_first, c_first+27) -> [N,4] (TRAIN:225-232)."""
    N, C, h, w = feat.shape
    out = torch.empty((N, 4), dtype=torch.float32, device=feat.data.device)
    rc = _lib.load().gfr_light_head_p16_fwd(_ptr(feat.data), feat.groups, int(c_first), h * w, _ptr(_need(w1, torch.float32, "w1")),
                                            _ptr(_need(b1, torch.float32, "b1")), _ptr(_need(w2, torch.float32, "w2")),
                                            _ptr(_need(b2, torch.float32, "b2")), _ptr(out), N, _stream())
    _lib.check(rc, "gfr_light_head_p16_fwd"); _count()
    return out


# ---------------------------------------------------------------------------------------------------
# output stage of the inference drivers (csrc/postprocess.cu): TEST1:590-620, TESTB:589-608, the MATLAB border fix
def _mask_u8(mask, B, H, W):
    if not (torch.is_tensor(mask) and mask.is_cuda and mask.dtype == torch.uint8):
        raise RuntimeError("mask must be a CUDA uint8 tensor holding the skin-mask values (0..255)")
    mask = mask.contiguous()
    if mask.numel() == H * W:
        return mask, 0
    if mask.numel() == B * H * W:
        return mask, H * W
    raise RuntimeError("mask must be [H,W] (shared) or [B,H,W]")


def composite_bgr_u8(image, rendered, mask):
    """image [B,H,W,3] RGB in [0,1] (CUDA f64 like the reference's `training_images`, or f32), rendered [B,3,H,W] f32,
    mask u8 [H,W] | [B,H,W] -> [B,H,W,3] u8 BGR: the array cv2.imwrite stores at TEST1:620."""
    rendered = _need(rendered, torch.float32, "rendered")
    if not (torch.is_tensor(image) and image.is_cuda and image.dtype in (torch.float32, torch.float64)):
        raise RuntimeError("image must be a CUDA f32/f64 tensor")
    image = image.contiguous()
    B, _, H, W = rendered.shape
    if tuple(image.shape) != (B, H, W, 3):
        raise RuntimeError("image must be [B,H,W,3]")
    mask, stride = _mask_u8(mask, B, H, W)
    out = torch.empty((B, H, W, 3), dtype=torch.uint8, device=rendered.device)
    rc = _lib.load().gfr_composite_bgr_u8(_ptr(image), int(image.dtype == torch.float64), _ptr(rendered), _ptr(mask), stride,
                                          _ptr(out), B, H, W, _stream())
    _lib.check(rc, "gfr_composite_bgr_u8"); _count()
    return out


def export_planes_u8(albedo, depth, shadow, final_shading, normals, mask,
                     want=("shadow_mask", "albedo", "depth", "shading", "surface_normals")):
    """TESTB:590-608: the five auxiliary images as u8 (BGR for the 3-channel ones).  Returns a dict."""
    depth = _need(depth, torch.float32, "depth")
    B, _, H, W = depth.shape
    mask, stride = _mask_u8(mask, B, H, W)
    dev = depth.device
    src = {"shadow_mask": shadow, "albedo": albedo, "depth": depth, "shading": final_shading, "surface_normals": normals}
    for k in want:
        src[k] = _need(src[k], torch.float32, k)
    out = {k: (torch.empty((B, H, W, 3) if k in ("albedo", "surface_normals") else (B, H, W), dtype=torch.uint8, device=dev)
               if k in want else None) for k in src}
    keys = None
    lib = _lib.load()
    if "depth" in want:
        keys = torch.empty(2, dtype=torch.int32, device=dev)
        rc = lib.gfr_neg_depth_range(_ptr(depth), depth.numel(), _ptr(keys), _stream())
        _lib.check(rc, "gfr_neg_depth_range"); _count(2)
    g = lambda k: _ptr(src[k]) if k in want else None
    rc = lib.gfr_export_planes_u8(g("albedo"), g("depth"), g("shadow_mask"), g("shading"), g("surface_normals"), _ptr(mask),
                                  stride, _ptr(keys), _ptr(out["shadow_mask"]), _ptr(out["albedo"]), _ptr(out["depth"]),
                                  _ptr(out["shading"]), _ptr(out["surface_normals"]), B, H, W, _stream())
    _lib.check(rc, "gfr_export_planes_u8"); _count()
    return {k: v for k, v in out.items() if v is not None}


def border_median_fix_u8(img, mask, max_sum=30):
    """fix_border_artifacts_CVPR2022.m on a u8 image batch [B,H,W,C] (or [B,H,W]); mask u8 [H,W] | [B,H,W]."""
    if not (torch.is_tensor(img) and img.is_cuda and img.dtype == torch.uint8):
        raise RuntimeError("img must be a CUDA uint8 tensor")
    img = img.contiguous()
    shape = img.shape
    B, H, W = shape[:3]
    C = shape[3] if img.dim() == 4 else 1
    mask, stride = _mask_u8(mask, B, H, W)
    out = torch.empty_like(img)
    rc = _lib.load().gfr_border_median_fix_u8(_ptr(img), _ptr(mask), stride, _ptr(out), B, H, W, C, int(max_sum), _stream())
    _lib.check(rc, "gfr_border_median_fix_u8"); _count()
    return out


def masked_mse_u8(recon, gt, mask):
    """MSE_MP.m:15-25: per-image masked MSE of two uint8 image batches [B,H,W,C]; mask u8 [H,W] | [B,H,W] -> [B] f64."""
    for t, n in ((recon, "recon"), (gt, "gt")):
        if not (torch.is_tensor(t) and t.is_cuda and t.dtype == torch.uint8 and t.dim() == 4):
            raise RuntimeError("%s must be a CUDA uint8 tensor [B,H,W,C]" % n)
    if recon.shape != gt.shape:
        raise RuntimeError("recon and gt must have the same shape")
    recon, gt = recon.contiguous(), gt.contiguous()
    B, H, W, C = recon.shape
    mask, stride = _mask_u8(mask, B, H, W)
    sums = torch.empty((B, 2), dtype=torch.float64, device=recon.device)
    rc = _lib.load().gfr_masked_mse_u8(_ptr(recon), _ptr(gt), _ptr(mask), stride, _ptr(sums), B, H, W, C, _stream())
    _lib.check(rc, "gfr_masked_mse_u8"); _count()
    return sums[:, 0] / (C * sums[:, 1])


def masked_dssim_u8(recon, gt, mask, window_3d=True):
    """DSSIM_MP_RGB.m:15-27: per-image masked DSSIM of two uint8 RGB image batches [B,H,W,3]; mask u8 [H,W] | [B,H,W]
    -> [B] f64 = (1 - sum(ssimmap * mask3) / sum(mask3)) / 2 with MATLAB's volume-window `ssim` map."""
    for t, n in ((recon, "recon"), (gt, "gt")):
        if not (torch.is_tensor(t) and t.is_cuda and t.dtype == torch.uint8 and t.dim() == 4 and t.shape[3] == 3):
            raise RuntimeError("%s must be a CUDA uint8 tensor [B,H,W,3]" % n)
    if recon.shape != gt.shape:
        raise RuntimeError("recon and gt must have the same shape")
    recon, gt = recon.contiguous(), gt.contiguous()
    B, H, W, _ = recon.shape
    mask, stride = _mask_u8(mask, B, H, W)
    sums = torch.empty((B, 2), dtype=torch.float64, device=recon.device)
    rc = _lib.load().gfr_masked_ssim_u8(_ptr(recon), _ptr(gt), _ptr(mask), stride, _ptr(sums), B, H, W, int(bool(window_3d)), _stream())
    _lib.check(rc, "gfr_masked_ssim_u8"); _count()
    return (1.0 - sums[:, 0] / sums[:, 1]) / 2.0
