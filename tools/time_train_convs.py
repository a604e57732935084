"""Dev tool: CUDA-event timing of the bf16 training convolutions (csrc/conv_tc.cu, precision 4) at PatchGAN's 2x2-tap layer shapes and
the generator's wide layers, B = 16: forward-direction launches replayed from a CUDA graph.  Prints us and TFLOP/s per layer."""
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
# (Cin of the operand, Cout, input H = W, taps): PatchGAN conv1..conv4 over the space-to-depth (TRAIN:18-27) and their data gradients,
# then 3x3 generator layers
shapes = [(12, 64, 129, 4, 0), (256, 128, 65, 4, 0), (512, 256, 33, 4, 0), (1024, 512, 17, 4, 0),
          (128, 256, 64, 4, 1), (256, 512, 32, 4, 1), (512, 1024, 16, 4, 1),
          (16, 16, 256, 9, 1), (32, 32, 64, 9, 1), (64, 64, 32, 9, 1), (155, 155, 16, 9, 1), (155, 128, 16, 9, 1)]
NCU = len(sys.argv) > 1 and sys.argv[1] == "ncu"        # eager launches of a few shapes (for an ncu capture)
if NCU:
    shapes = [shapes[i] for i in (0, 1, 3, 4, 7)]
for Cin, Cout, S, taps, org in shapes:
    G = (Cin + 3) // 4
    x = torch.randn(B, G, S, S, 4, device="cuda")
    w = torch.randn(Cout, Cin, 2 if taps == 4 else 3, 2 if taps == 4 else 3, device="cuda") * 0.05
    b = torch.zeros(Cout, device="cuda")
    NT = train_ops._nt_for(Cout, taps, 4)
    packed = train_ops._pack_dev(w, False, False, Cin, Cout, NT, taps, 4)
    if NCU:
        for _ in range(2): train_ops._conv_raw(x, Cin, packed, b, Cout, NT, taps, org, 4)
        torch.cuda.synchronize()
        continue
    us = t(lambda: train_ops._conv_raw(x, Cin, packed, b, Cout, NT, taps, org, 4))
    Ho = S if taps == 9 else (S - 1 if org == 0 else S + 1)
    flop = 2.0 * B * Ho * Ho * Cin * Cout * taps
    us_p = t(lambda: train_ops._pack_dev(w, False, False, Cin, Cout, NT, taps, 4))
    print("%4d->%4d in %3d^2 taps %d NT %3d: %7.1f us  %6.1f TFLOP/s   (pack %5.1f us)" % (Cin, Cout, S, taps, NT, us, flop / us / 1e6, us_p))

if not NCU:
    print("weight gradients (gfr_conv_wgrad_tc_bf16), B = 16:")
    for Cin, Cout, S, taps, deconv in [(16, 16, 256, 9, False), (16, 16, 128, 9, False), (32, 32, 64, 9, False), (64, 64, 32, 9, False), (64, 64, 32, 9, True),
                                       (155, 155, 16, 9, False), (128, 155, 16, 9, True), (155, 64, 32, 9, True), (256, 128, 64, 4, False), (1024, 512, 16, 4, False)]:
        G = (Cin + 3) // 4
        Hin = S + 1 if taps == 4 else S
        x = torch.randn(B, G, Hin, Hin, 4, device="cuda")
        g = torch.randn(B, (Cout + 3) // 4, S, S, 4, device="cuda")
        k = 2 if taps == 4 else 3
        w = torch.zeros((Cin, Cout, k, k) if deconv else (Cout, Cin, k, k), device="cuda", requires_grad=True)
        w.grad = torch.zeros_like(w)
        us = t(lambda: train_ops._wgrad(x, g, w, None, deconv, Cin, Cout, taps, 4))
        flop = 2.0 * B * S * S * Cin * Cout * taps
        print("%4d->%4d @%3d^2 taps %d %s: %7.1f us  %6.1f TFLOP/s" % (Cin, Cout, S, taps, "deconv" if deconv else "conv  ", us, flop / us / 1e6))
