"""Dev tool: CUDA-event timing of the tcgen05 conv against the fp32 direct conv at the RelightNet layer shapes."""
import sys, os, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geomconsistentfr_b200 import ops

def t(fn, n=20):
    """n back-to-back launches replayed from a CUDA graph (no CPU launch cost in the number)."""
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        for _ in range(3): fn()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(n): fn()
        g.replay(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(s)
        g.replay()
        e1.record(s); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3

B = 8
if len(sys.argv) > 1 and sys.argv[1] == "one":      # a single configuration, eager launches (for ncu)
    Cin, Cout, S, NT = [int(v) for v in sys.argv[2:6]]
    x = torch.randn(B, Cin, S, S, device="cuda"); w = torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.05; b = torch.zeros(Cout, device="cuda")
    xc = ops.nchw_to_c4(x); wp = ops.conv_tc_pack_weights(w, NT)
    for _ in range(5): ops.conv3x3_tc_fwd(xc, wp, b, Cout, NT)
    torch.cuda.synchronize(); sys.exit(0)
for (Cin, Cout, S, NT) in [(16, 16, 16, 16), (16, 16, 256, 16), (16, 16, 128, 16), (32, 32, 64, 32), (32, 32, 64, 16), (64, 64, 32, 64), (64, 64, 32, 32), (64, 64, 32, 16),
                           (155, 155, 16, 32), (155, 155, 16, 16), (128, 64, 16, 16), (64, 155, 16, 16)]:
    x = torch.randn(B, Cin, S, S, device="cuda"); w = torch.randn(Cout, Cin, 3, 3, device="cuda") * 0.05; b = torch.zeros(Cout, device="cuda")
    xc = ops.nchw_to_c4(x); wp = ops.conv_tc_pack_weights(w, NT)
    us_tc = t(lambda: ops.conv3x3_tc_fwd(xc, wp, b, Cout, NT))
    us_1 = t(lambda: ops.conv3x3_tc_fwd(xc, wp, b, Cout, NT, precision=1))
    wp16, ws16, nt16 = ops.conv_tc_pack_weights_f16(w)
    us_16 = t(lambda: ops.conv3x3_tc_fwd(xc, wp16, b, Cout, nt16, precision=2, w_scale=ws16))
    us_d = t(lambda: ops.conv2d_fwd(x, w, b))
    mac = B * S * S * Cin * Cout * 9
    print("%3d->%3d @%3d NT=%2d: tc %8.1f us (%6.2f TMAC/s)   1xTF32 %8.1f us   fp16x2(NT%d) %8.1f us   direct %8.1f us" % (Cin, Cout, S, NT, us_tc, mac / us_tc / 1e6, us_1, nt16, us_16, us_d))
