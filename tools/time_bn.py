"""Dev tool: CUDA-event timing of the train-mode BatchNorm passes (csrc/cnn_train.cu) at the generator's layer shapes, B = 16:
statistics, apply, backward (reduce + apply) — us and the HBM traffic rate of each (tensors > L2 at 256^2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from geomconsistentfr_b200 import train_ops


def t(fn, n=10):
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(n): fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s); g.replay(); e1.record(s); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


B = 16
for C, S, res in [(16, 256, False), (16, 256, True), (16, 128, True), (32, 64, True), (64, 32, False), (155, 16, True)]:
    G = (C + 3) // 4
    bn = torch.nn.BatchNorm2d(C).cuda().train()
    raw = torch.randn(B, G, S, S, 4, device="cuda")
    r = torch.randn_like(raw) if res else None
    gy = torch.randn_like(raw)
    mean, rstd, scale, shift = train_ops._BN.stats(raw, C, bn)
    mb = raw.numel() * 4 / 1e6
    us_s = t(lambda: train_ops._BN.stats(raw, C, bn))
    us_a = t(lambda: train_ops._BN.apply(raw, C, scale, shift, r, None, 0, 1))
    us_b = t(lambda: train_ops._BN.backward(raw, C, r, gy, mean, rstd, scale, shift, bn.weight, 1, res, bn.bias, None))
    n_a, n_b = (3 if res else 2), (2 * (3 if res else 2) + (2 if res else 1))
    print("C %3d @%3d^2 res %d (%5.1f MB/tensor): stats %6.1f us %5.2f TB/s | apply %6.1f us %5.2f TB/s | bwd %6.1f us %5.2f TB/s" %
          (C, S, res, mb, us_s, mb / us_s, us_a, n_a * mb / us_a, us_b, n_b * mb / us_b))

# the decoders' 1x1 tails (16 -> 16 at 256^2): forward and backward (data + weight gradient)
from geomconsistentfr_b200 import _lib
from geomconsistentfr_b200.ops import _ptr, _stream
x = torch.randn(B, 4, 256, 256, 4, device="cuda"); gy = torch.randn_like(x); y = torch.empty_like(x); gx = torch.empty_like(x)
w = torch.randn(16, 16, device="cuda") * 0.1; b = torch.zeros(16, device="cuda"); gw = torch.zeros_like(w); gb = torch.zeros_like(b)
L = _lib.load()
us_f = t(lambda: L.gfr_pw_conv16_fwd(_ptr(x), _ptr(w), _ptr(b), _ptr(y), B, 16, 256, 256, 0, 0, 1.0, _stream()))
us_b = t(lambda: L.gfr_pw_conv16_bwd(_ptr(x), _ptr(w), _ptr(gy), None, _ptr(gx), _ptr(gw), _ptr(gb), B, 16, 256, 256, 0, 0, 1.0, _stream()))
mb = x.numel() * 4 / 1e6
print("1x1 16->16 @256^2: fwd %6.1f us %5.2f TB/s | bwd %6.1f us %5.2f TB/s" % (us_f, 2 * mb / us_f, us_b, 3 * mb / us_b))
