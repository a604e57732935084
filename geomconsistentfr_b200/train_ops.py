"""Train-mode units of the RelightNet CNN as torch.autograd.Functions over libgfr_b200 kernels (C4 activations).

The reference trains with batch-statistics BatchNorm (it never calls .eval(), TRAIN:561-563), so conv and BN cannot be
folded: a unit is  conv (tcgen05, raw output) -> batch statistics -> y = act(BN(raw) + res) + up(post)  and its
backward is  BN/activation backward (two passes) -> dgrad (the same tcgen05 kernel with the transposed + flipped
operand) + wgrad (CUDA cores).  TRAIN = train_raytracing_relighting_CelebAHQ_DSSIM_8x.py, lines 197-350."""
import ctypes

import torch

from . import _lib, ops
from .ops import C4, _ptr, _stream

_EPS, _MOMENTUM = 1e-5, 0.1          # nn.BatchNorm2d defaults (TRAIN:59 ff.)


def _chk(rc, what, n=1):
    _lib.check(rc, what)
    ops._count(n)


class PackPlan:
    """The packed conv operands of a training step as ONE launch (gfr_conv_tc_pack_weights_batch).

    A step re-packs every layer's forward and data-gradient operand after the optimiser moved the parameters (TRAIN:656):
    ~120 launches of 2-4 us, each in front of the layer that needs it.  While a plan is active (`with plan:`) `_pack_dev`
    registers every (parameter storage, direction, tile) it is asked for — parameters only: a temporary (PatchGAN's permuted
    weight view) has no stable address — and hands out a persistent buffer; `plan.run()` at the start of the next step fills
    all registered buffers in one launch, and `_pack_dev` then returns them without launching.  `plan.invalidate()` after
    an optimiser step.  The job table is rebuilt (one small host->device copy) when new jobs appeared — never during a
    stream capture: the eager warm-up steps in front of a capture register them."""

    active = None

    def __init__(self, params=()):
        self.param_ptrs = {p.data_ptr() for p in params}     # storages with a stable address (call after FlatAdam re-seated them)
        self.jobs = {}                 # key -> [w, deconv, dgrad, Cin, Cout, NT, taps, precision, packed, in the device table?]
        self.fresh = False
        self._table = None
        self._n_table = 0
        self._blocks = 0

    def __enter__(self):
        self._outer, PackPlan.active = PackPlan.active, self
        return self

    def __exit__(self, *exc):
        PackPlan.active = self._outer
        return False

    def invalidate(self):
        self.fresh = False

    def run(self):
        """Pack every registered operand from the current parameter values (one launch)."""
        if not self.jobs:
            return
        if self._n_table != len(self.jobs):
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("PackPlan: new conv layers appeared during a stream capture (run an eager step first)")
            lib = _lib.load()
            rec = lib.gfr_conv_tc_pack_job_size()
            buf = ctypes.create_string_buffer(rec * len(self.jobs))
            first = 0
            for i, (w, deconv, dgrad, Cin, Cout, NT, taps, prec, packed, _) in enumerate(self.jobs.values()):
                n = lib.gfr_conv_tc_pack_job_fill(ctypes.c_void_p(ctypes.addressof(buf) + i * rec), _ptr(w), int(deconv), int(dgrad), Cin, Cout,
                                                  NT, taps, prec, _ptr(packed), first)
                if n < 0:
                    _lib.check(int(n), "gfr_conv_tc_pack_job_fill")
                first += n
            dev = next(iter(self.jobs.values()))[0].device
            self._table = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).to(dev)
            self._n_table, self._blocks = len(self.jobs), first
            for job in self.jobs.values():
                job[9] = True                                  # in the table: filled by every run() from now on
        _chk(_lib.load().gfr_conv_tc_pack_weights_batch(_ptr(self._table), self._n_table, self._blocks, _stream()),
             "gfr_conv_tc_pack_weights_batch")
        self.fresh = True

    def get(self, w, deconv, dgrad, Cin, Cout, NT, taps, precision):
        """-> (packed buffer, already filled by run()?) for a parameter; None for a tensor that is not one."""
        if w.data_ptr() not in self.param_ptrs:
            return None
        key = (w.data_ptr(), int(deconv), int(dgrad), Cin, Cout, NT, taps, precision)
        job = self.jobs.get(key)
        if job is None:
            O, I = (Cin, Cout) if dgrad else (Cout, Cin)
            n = _lib.load().gfr_conv_tc_pack_size_ex(I, O, NT, taps, precision)
            job = self.jobs[key] = [w, int(deconv), int(dgrad), Cin, Cout, NT, taps, precision,
                                    torch.empty(n, dtype=torch.float32, device=w.device), False]
        return job[8], self.fresh and job[9]


def _pack_dev(w, deconv, dgrad, Cin, Cout, NT, taps=9, precision=3):
    """Device-side packing of the layer parameter for the tensor-core conv (forward or data-gradient operand)."""
    plan = PackPlan.active
    if plan is not None:
        got = plan.get(w, deconv, dgrad, Cin, Cout, NT, taps, precision)
        if got is not None:
            packed, filled = got
            if not filled:
                _chk(_lib.load().gfr_conv_tc_pack_weights_dev_ex(_ptr(w), int(deconv), int(dgrad), Cin, Cout, NT, taps, precision, _ptr(packed),
                                                                 _stream()), "gfr_conv_tc_pack_weights_dev_ex")
            return packed
    O, I = (Cin, Cout) if dgrad else (Cout, Cin)
    n = _lib.load().gfr_conv_tc_pack_size_ex(I, O, NT, taps, precision)
    packed = torch.empty(n, dtype=torch.float32, device=w.device)
    _chk(_lib.load().gfr_conv_tc_pack_weights_dev_ex(_ptr(w), int(deconv), int(dgrad), Cin, Cout, NT, taps, precision, _ptr(packed),
                                                     _stream()), "gfr_conv_tc_pack_weights_dev_ex")
    return packed


def _nt_for(cout, taps=9, precision=3):
    """Output channels per CTA tile.  bf16 accumulates in TMEM and reads it in 16-column groups, so wide tiles (up to 128)
    amortise the A-operand reads; the fp32-grade kinds keep per-step register accumulators (<= 64)."""
    if precision == 4:
        if taps == 4:
            return 16 if cout <= 16 else (64 if cout <= 64 else 128)
        return 16 if cout <= 16 else (32 if cout <= 32 else (64 if cout <= 64 else 128))
    if taps == 4:
        return 16 if cout <= 16 else 64
    return 16 if cout <= 16 else 32


def _conv_raw(x_data, cin, packed, bias, Cout, NT, taps=9, org=1, precision=3, act=0):
    """raw = conv(x[:, :cin]) + bias on the tensor cores, C4 in / C4 out.  taps 9: 3x3 / pad 1; taps 4: the 2x2-tap layers
    (org 0: (H+1)x(W+1) -> HxW, org 1: HxW -> (H+1)x(W+1))."""
    N, G, Hin, Win, _ = x_data.shape
    H, W = (Hin, Win) if taps == 9 else ((Hin - 1, Win - 1) if org == 0 else (Hin + 1, Win + 1))
    out = torch.empty((N, (Cout + 3) // 4, H, W, 4), dtype=torch.float32, device=x_data.device)
    _chk(_lib.load().gfr_conv_tc_fwd_ex(_ptr(x_data), _ptr(packed), _ptr(bias), None, None, _ptr(out), N, cin, G, Cout, Hin, Win, H, W,
                                        NT, taps, org, 0, int(act), 1.0, int(precision), 0, _stream()), "gfr_conv_tc_fwd_ex")
    return out


def _grad_target(p, like=None):
    """Where a parameter gradient is accumulated.  A leaf parameter whose .grad already exists (FlatAdam keeps every .grad as a
    view of ONE flat buffer, zeroed once per step) is accumulated IN PLACE by the kernels (they all `+=`), and the autograd
    Function returns None for it: no zeros() temporary, no fill, no AccumulateGrad add per parameter — ~600 tiny launches per
    training iteration.  Otherwise a fresh zero tensor that autograd accumulates as usual.  -> (tensor, direct)"""
    if p is not None and p.is_leaf and p.requires_grad and p.grad is not None and p.grad.is_contiguous() and p.grad.dtype == torch.float32:
        return p.grad, True
    ref = p if like is None else like
    return torch.zeros(ref.shape, dtype=torch.float32, device=ref.device), False


_ZEROS = {}


def _zeros(n, device):
    """A cached all-zero vector (the bias operand of the data-gradient convolutions)."""
    z = _ZEROS.get(device)
    if z is None or z.numel() < n:
        z = _ZEROS[device] = torch.zeros(max(n, 2048), dtype=torch.float32, device=device)
    return z


def _wgrad(x, g_raw, w, b, deconv, cin, Cout, taps=9, precision=3):
    """Weight / bias gradient of the conv layer from its input and dL/d(conv output) -> (g_w, g_b) to return from the autograd
    Function (None where the gradient went straight into the parameter's .grad).  precision 4: bf16 on the tensor cores
    (csrc/wgrad_tc.cu); otherwise fp32 on CUDA cores."""
    N, G = x.shape[0], x.shape[1]
    H, W = g_raw.shape[2], g_raw.shape[3]
    g_w, dw = _grad_target(w)
    g_b, db = _grad_target(b) if (b is not None and b.requires_grad) else (None, True)
    if precision == 4:
        _chk(_lib.load().gfr_conv_wgrad_tc_bf16(_ptr(x), _ptr(g_raw), _ptr(g_w), int(deconv), N, cin, G, Cout, x.shape[2], x.shape[3],
                                                H, W, taps, _stream()), "gfr_conv_wgrad_tc_bf16")
        if g_b is not None:
            _chk(_lib.load().gfr_channel_sum_c4(_ptr(g_raw), _ptr(g_b), N, Cout, H, W, _stream()), "gfr_channel_sum_c4")
    elif taps == 9:
        _chk(_lib.load().gfr_conv3x3_wgrad(_ptr(x), _ptr(g_raw), _ptr(g_w), _ptr(g_b), int(deconv), N, cin, G, Cout, H, W, _stream()),
             "gfr_conv3x3_wgrad", 2)
    else:
        _chk(_lib.load().gfr_conv2x2_wgrad(_ptr(x), _ptr(g_raw), _ptr(g_w), _ptr(g_b), N, cin, G, Cout, H, W, _stream()),
             "gfr_conv2x2_wgrad", 2)
    return (None if dw else g_w), (None if db else g_b)


class ZeroArena:
    """The fp64 scratch of every BatchNorm pass of a training step (batch sums, ticket counters) as slices of ONE buffer that
    `begin()` clears with a single memset at the start of the step; while the arena is active the library skips the memset node
    it otherwise puts in front of each of the ~150 BatchNorm passes of an iteration (gfr_bn_config)."""

    active = None

    def __init__(self, n_doubles=1 << 17):
        self.n, self.buf, self.off, self.high = n_doubles, None, 0, 0

    def begin(self, device):
        if self.buf is None:
            self.buf = torch.zeros(self.n, dtype=torch.float64, device=device)
        elif self.high:
            self.buf[:self.high].zero_()
        self.off = 0

    def take(self, n):
        if self.buf is None:
            return None
        end = self.off + ((n + 1) & ~1)                   # 16-byte aligned slices
        if end > self.n:
            return torch.zeros(n, dtype=torch.float64, device=self.buf.device)      # arena exhausted: a zeroed temporary
        v = self.buf[self.off:self.off + n]
        self.off = end
        self.high = max(self.high, end)
        return v

    def __enter__(self):
        self._outer = ZeroArena.active
        ZeroArena.active = self
        self._old = _lib.load().gfr_bn_config(1)
        return self

    def __exit__(self, *exc):
        ZeroArena.active = self._outer
        _lib.load().gfr_bn_config(self._old)
        return False


def _bn_scratch(n, device):
    a = ZeroArena.active
    if a is not None:
        v = a.take(n)
        if v is not None:
            return v
        return torch.zeros(n, dtype=torch.float64, device=device)
    return torch.empty(n, dtype=torch.float64, device=device)


class _BN:
    """Batch statistics + apply + backward of one BatchNorm2d over a C4 tensor (shared by the unit Functions)."""

    @staticmethod
    def stats(raw, C, bn):
        N, G, H, W, _ = raw.shape
        dev = raw.device
        sums = _bn_scratch(2 * G * 4 + 1, dev)          # + the ticket counter of the fused finalise
        mean, rstd, scale, shift = (torch.empty(G * 4, dtype=torch.float32, device=dev) for _ in range(4))
        track = bn.training and bn.track_running_stats
        nbt = bn.num_batches_tracked if (track and bn.num_batches_tracked is not None and bn.num_batches_tracked.is_cuda
                                         and bn.num_batches_tracked.dtype == torch.int64) else None
        _chk(_lib.load().gfr_bn_train_stats_ex(_ptr(raw), _ptr(bn.weight), _ptr(bn.bias), _ptr(bn.running_mean) if track else None,
                                               _ptr(bn.running_var) if track else None, _ptr(nbt), _ptr(sums), _ptr(mean), _ptr(rstd),
                                               _ptr(scale), _ptr(shift), N, C, H, W, float(bn.eps), float(bn.momentum), _stream()),
             "gfr_bn_train_stats_ex", 1)
        if track:
            if nbt is None and bn.num_batches_tracked is not None:
                bn.num_batches_tracked += 1          # (a counter that is not a CUDA int64 tensor: torch's own increment)
            ops.bump_param_generation()
            if _BN.recording is not None:
                _BN.recording.append((bn, sums, N, C, H, W))
        return mean, rstd, scale, shift

    recording = None          # a list while `record_running_updates()` is active: (bn, batch sums, N, C, H, W) of every tracked BatchNorm pass

    @staticmethod
    def replay_running_updates(records):
        """Apply the running-statistics update of the recorded passes once more (gfr_bn_running_update) — what a repeated forward
        over the same input with the same weights does to the BatchNorm buffers."""
        for bn, sums, N, C, H, W in records:
            nbt = bn.num_batches_tracked if (bn.num_batches_tracked is not None and bn.num_batches_tracked.is_cuda
                                             and bn.num_batches_tracked.dtype == torch.int64) else None
            _chk(_lib.load().gfr_bn_running_update(_ptr(sums), _ptr(bn.running_mean), _ptr(bn.running_var), _ptr(nbt), N, C, H, W,
                                                   float(bn.momentum), _stream()), "gfr_bn_running_update")
            if nbt is None and bn.num_batches_tracked is not None:
                bn.num_batches_tracked += 1
        if records:
            ops.bump_param_generation()

    @staticmethod
    def apply(raw, C, scale, shift, res, post, post_shift, act):
        N, G, H, W, _ = raw.shape
        y = torch.empty_like(raw)
        _chk(_lib.load().gfr_bn_apply_fwd(_ptr(raw), _ptr(scale), _ptr(shift), _ptr(res), _ptr(post), _ptr(y), N, C, H, W,
                                          int(post_shift), int(act), _stream()), "gfr_bn_apply_fwd")
        return y

    @staticmethod
    def backward(raw, C, res, g_y, mean, rstd, scale, shift, gamma, act, want_res, beta=None, conv_bias=None):
        """-> g_raw, g_res, g_gamma, g_beta, g_conv_bias (None where a gradient was accumulated straight into the parameter's
        .grad; g_conv_bias = sum of g_raw per channel, from the same pass, when `conv_bias` is given)"""
        N, G, H, W, _ = raw.shape
        dev = raw.device
        sums = _bn_scratch(2 * G * 4, dev)
        g_raw = torch.empty_like(raw)
        g_res = torch.empty_like(raw) if want_res else None
        g_gamma, dg = _grad_target(gamma)
        g_beta, db = _grad_target(beta, like=gamma)
        g_cb, dcb = (_grad_target(conv_bias) if conv_bias is not None and conv_bias.requires_grad else (None, True))
        _chk(_lib.load().gfr_bn_apply_bwd_ex(_ptr(raw), _ptr(res), _ptr(g_y), _ptr(scale), _ptr(shift), _ptr(mean), _ptr(rstd),
                                             _ptr(gamma.detach().contiguous()), _ptr(sums), _ptr(g_raw), _ptr(g_res), _ptr(g_gamma),
                                             _ptr(g_beta), _ptr(g_cb), N, C, H, W, int(act), _stream()), "gfr_bn_apply_bwd_ex", 2)
        return g_raw, g_res, (None if dg else g_gamma), (None if db else g_beta), (None if dcb else g_cb)


class record_running_updates:
    """Context manager: collects the batch statistics of every tracked BatchNorm forward inside it -> `.records` for
    `_BN.replay_running_updates`."""

    def __enter__(self):
        self.records, self._outer = [], _BN.recording
        _BN.recording = self.records
        return self

    def __exit__(self, *exc):
        _BN.recording = self._outer
        return False


def _sumpool2(g):
    N, G, H, W, _ = g.shape
    out = torch.empty((N, G, H // 2, W // 2, 4), dtype=torch.float32, device=g.device)
    _chk(_lib.load().gfr_sumpool2_c4(_ptr(g), _ptr(out), N * G, H // 2, W // 2, _stream()), "gfr_sumpool2_c4")
    return out


class ConvBNAct(torch.autograd.Function):
    """y = act(BN_train(conv3x3(x[:, :cin]) + b) + res) + up(post).  Tensors are C4 `data` arrays."""

    @staticmethod
    def forward(ctx, x, w, b, gamma, beta, res, post, meta):
        cin, Cout, deconv, act, post_shift, bn = meta["cin"], meta["cout"], meta["deconv"], meta["act"], meta["post_shift"], meta["bn"]
        taps, prec = meta.get("taps", 9), meta.get("precision", 3)
        NT = _nt_for(Cout, taps, prec)
        raw = _conv_raw(x, cin, _pack_dev(w, deconv, False, cin, Cout, NT, taps, prec), b, Cout, NT, taps, 0 if taps == 4 else 1, prec)
        mean, rstd, scale, shift = _BN.stats(raw, Cout, bn)
        y = _BN.apply(raw, Cout, scale, shift, res, post, post_shift, act)
        ctx.save_for_backward(x, w, raw, res, mean, rstd, scale, shift, gamma, beta, b)
        ctx.meta = meta
        ctx.has = (res is not None, post is not None)
        return y

    @staticmethod
    def backward(ctx, g_y):
        x, w, raw, res, mean, rstd, scale, shift, gamma, beta, b = ctx.saved_tensors
        m = ctx.meta
        cin, Cout, deconv, act, post_shift = m["cin"], m["cout"], m["deconv"], m["act"], m["post_shift"]
        taps, prec = m.get("taps", 9), m.get("precision", 3)
        has_res, has_post = ctx.has
        g_y = g_y.contiguous()
        N, G, H, W, _ = x.shape
        g_post = None
        if has_post:
            g_post = _sumpool2(g_y) if post_shift else g_y
        g_raw, g_res, g_gamma, g_beta, g_b = _BN.backward(raw, Cout, res, g_y, mean, rstd, scale, shift, gamma, act, has_res, beta, b)
        # data gradient: the same tensor-core convolution, Cout -> cin, transposed + flipped kernel
        g_x = None
        if ctx.needs_input_grad[0]:
            NTd = _nt_for(cin, taps, prec)
            g_in = _conv_raw(g_raw, Cout, _pack_dev(w, deconv, True, cin, Cout, NTd, taps, prec), _zeros(cin, x.device), cin, NTd, taps, 1, prec)
            if g_in.shape[1] == G:
                g_x = g_in
            else:                                   # the layer read only the leading channels of a wider tensor (TRAIN:225)
                g_x = torch.zeros_like(x)
                g_x[:, :g_in.shape[1]] = g_in
        g_w = None
        if ctx.needs_input_grad[1]:          # frozen weights (the generator's pass through the discriminator) skip the wgrad
            g_w, _ = _wgrad(x, g_raw, w, None, deconv, cin, Cout, taps, prec)        # (the bias gradient came out of the BN backward)
        else:
            g_b = None
        return g_x, g_w, g_b, g_gamma, g_beta, g_res, g_post, None


class StemBNAct(torch.autograd.Function):
    """c1_og = LeakyReLU(BN_train(conv5x5(img) + b)) on the NHWC image (TRAIN:197-200).  No gradient w.r.t. the image."""

    @staticmethod
    def forward(ctx, img, w, b, gamma, beta, bn):
        N, H, W, _ = img.shape
        raw = torch.empty((N, 4, H, W, 4), dtype=torch.float32, device=img.device)
        _chk(_lib.load().gfr_stem_conv_train_fwd(_ptr(img), _ptr(w), _ptr(b), _ptr(raw), N, H, W, _stream()), "gfr_stem_conv_train_fwd")
        mean, rstd, scale, shift = _BN.stats(raw, 16, bn)
        y = _BN.apply(raw, 16, scale, shift, None, None, 0, 1)
        ctx.save_for_backward(img, raw, mean, rstd, scale, shift, gamma, beta, w, b)
        return y

    @staticmethod
    def backward(ctx, g_y):
        img, raw, mean, rstd, scale, shift, gamma, beta, w, b = ctx.saved_tensors
        N, H, W, _ = img.shape
        g_raw, _, g_gamma, g_beta, _ = _BN.backward(raw, 16, None, g_y.contiguous(), mean, rstd, scale, shift, gamma, 1, False, beta)
        (g_w, dw), (g_b, db) = _grad_target(w), _grad_target(b)
        _chk(_lib.load().gfr_stem_conv_wgrad(_ptr(img), _ptr(g_raw), _ptr(g_w), _ptr(g_b), N, H, W, _stream()), "gfr_stem_conv_wgrad")
        return None, (None if dw else g_w), (None if db else g_b), g_gamma, g_beta, None


class MaxPool2(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        N, G, H, W, _ = x.shape
        y = torch.empty((N, G, H // 2, W // 2, 4), dtype=torch.float32, device=x.device)
        _chk(_lib.load().gfr_maxpool2_c4_fwd(_ptr(x), _ptr(y), N * G, H // 2, W // 2, _stream()), "gfr_maxpool2_c4_fwd")
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, g_y):
        (x,) = ctx.saved_tensors
        N, G, H, W, _ = x.shape
        g_x = torch.empty_like(x)
        _chk(_lib.load().gfr_maxpool2_c4_bwd(_ptr(x), _ptr(g_y.contiguous()), _ptr(g_x), N * G, H // 2, W // 2, _stream()),
             "gfr_maxpool2_c4_bwd")
        return g_x


class Upsample2(torch.autograd.Function):
    """nearest x2 (TRAIN:240 ...) when the epoch gate keeps the skip branch off."""

    @staticmethod
    def forward(ctx, x):
        N, G, H, W, _ = x.shape
        y = torch.empty((N, G, 2 * H, 2 * W, 4), dtype=torch.float32, device=x.device)
        _chk(_lib.load().gfr_upsample2_c4_fwd(_ptr(x), None, _ptr(y), N * G, 2 * H, 2 * W, _stream()), "gfr_upsample2_c4_fwd")
        return y

    @staticmethod
    def backward(ctx, g_y):
        return _sumpool2(g_y.contiguous())


class PwConvBNAct(torch.autograd.Function):
    """y = LeakyReLU(BN_train(conv1x1(x) + b)), 16 -> 16 (c2_2, c2_3; TRAIN:286-287)."""

    @staticmethod
    def forward(ctx, x, w, b, gamma, beta, bn):
        N, G, H, W, _ = x.shape
        raw = torch.empty_like(x)
        _chk(_lib.load().gfr_pw_conv16_fwd(_ptr(x), _ptr(w), _ptr(b), _ptr(raw), N, 16, H, W, 0, 0, 1.0, _stream()), "gfr_pw_conv16_fwd")
        mean, rstd, scale, shift = _BN.stats(raw, 16, bn)
        y = _BN.apply(raw, 16, scale, shift, None, None, 0, 1)
        ctx.save_for_backward(x, w, raw, mean, rstd, scale, shift, gamma, beta, b)
        return y

    @staticmethod
    def backward(ctx, g_y):
        x, w, raw, mean, rstd, scale, shift, gamma, beta, b = ctx.saved_tensors
        N, G, H, W, _ = x.shape
        g_raw, _, g_gamma, g_beta, _ = _BN.backward(raw, 16, None, g_y.contiguous(), mean, rstd, scale, shift, gamma, 1, False, beta)
        g_x = torch.empty_like(x)
        g_w = torch.zeros_like(w)                    # (w is a .view() of the parameter: not a leaf, autograd routes it)
        (g_b, db) = _grad_target(b)
        _chk(_lib.load().gfr_pw_conv16_bwd(_ptr(x), _ptr(w), _ptr(g_raw), None, _ptr(g_x), _ptr(g_w), _ptr(g_b), N, 16, H, W, 0, 0,
                                           1.0, _stream()), "gfr_pw_conv16_bwd")
        return g_x, g_w, (None if db else g_b), g_gamma, g_beta, None


class PwHead(torch.autograd.Function):
    """out = scale * act(conv1x1(x) + b), 16 -> n_out, NCHW planes (c2_o + sigmoid / x100; TRAIN:289-290, 349-350)."""

    @staticmethod
    def forward(ctx, x, w, b, act, scale):
        N, G, H, W, _ = x.shape
        n_out = w.shape[0]
        out = torch.empty((N, n_out, H, W), dtype=torch.float32, device=x.device)
        _chk(_lib.load().gfr_pw_conv16_fwd(_ptr(x), _ptr(w), _ptr(b), _ptr(out), N, n_out, H, W, 1, act, float(scale), _stream()),
             "gfr_pw_conv16_fwd")
        ctx.save_for_backward(x, w, out)
        ctx.cfg = (act, scale)
        return out

    @staticmethod
    def backward(ctx, g_out):
        x, w, out = ctx.saved_tensors
        act, scale = ctx.cfg
        N, G, H, W, _ = x.shape
        n_out = w.shape[0]
        g_x = torch.empty_like(x)
        g_w = torch.zeros_like(w)
        g_b = torch.zeros(n_out, dtype=torch.float32, device=x.device)
        _chk(_lib.load().gfr_pw_conv16_bwd(_ptr(x), _ptr(w), _ptr(g_out.contiguous()), _ptr(out), _ptr(g_x), _ptr(g_w), _ptr(g_b), N,
                                           n_out, H, W, 1, act, float(scale), _stream()), "gfr_pw_conv16_bwd")
        return g_x, g_w, g_b, None, None


class AvgPoolChannels(torch.autograd.Function):
    """[N, n_ch] = mean over (H,W) of channels [c_first, c_first + n_ch) of a C4 map (TRAIN:226-230)."""

    @staticmethod
    def forward(ctx, x, C, c_first, n_ch):
        N, G, H, W, _ = x.shape
        out = torch.empty((N, n_ch), dtype=torch.float32, device=x.device)
        _chk(_lib.load().gfr_avgpool_c4_fwd(_ptr(x), _ptr(out), N, C, c_first, n_ch, H * W, _stream()), "gfr_avgpool_c4_fwd")
        ctx.cfg = (x.shape, C, c_first, n_ch)
        return out

    @staticmethod
    def backward(ctx, g):
        shape, C, c_first, n_ch = ctx.cfg
        N, G, H, W, _ = shape
        g_x = torch.zeros(shape, dtype=torch.float32, device=g.device)
        _chk(_lib.load().gfr_avgpool_c4_bwd(_ptr(g.contiguous()), _ptr(g_x), N, C, c_first, n_ch, H * W, _stream()), "gfr_avgpool_c4_bwd")
        return g_x, None, None, None
