"""Times the P16 conv kernel against the first-generation fp16-split kernel on the layer shapes of the B = 8 forward."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from geomconsistentfr_b200 import ops

SHAPES = [(8, 16, 16, 256), (8, 16, 16, 128), (8, 32, 32, 64), (8, 32, 16, 64), (8, 64, 64, 32), (8, 64, 128, 32), (8, 155, 155, 16),
          (8, 64, 320, 16), (8, 128, 128, 16)]


def timeit(fn, n=20):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        fn()
    ev = []
    torch.cuda._sleep(20_000_000)
    for _ in range(n):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record()
        ev.append((a, b))
    torch.cuda.synchronize()
    t = sorted(x.elapsed_time(y) for x, y in ev)
    return 1e3 * t[len(t) // 2]


for N, Cin, Cout, S in SHAPES:
    x = torch.randn(N, Cin, S, S, device="cuda")
    w = torch.randn(Cout, Cin, 3, 3, device="cuda") / (3 * Cin ** 0.5)
    b = torch.zeros(Cout, device="cuda")
    row = "%3d->%3d @%3d^2:" % (Cin, Cout, S)
    if Cout <= 160:
        x4 = ops.nchw_to_c4(x)
        wp, ws, NT = ops.conv_tc_pack_weights_f16(w)
        row += "  gen1 %6.1f us" % timeit(lambda: ops.conv3x3_tc_fwd(x4, wp, b, Cout, NT, precision=2, w_scale=ws))
    xp = ops.nchw_to_p16(x)
    for cfg in [(nt, mh, ks) for nt in (16, 32) for mh in (1, 2) for ks in (2, 4)]:
        if (cfg[0] == 16 and Cout > 16) or (cfg[2] == 4 and Cin < 32) or (cfg[0] == 32 and Cout <= 16):
            continue
        wq, s = ops.conv_p16_pack_weights(w, cfg[0], cfg[2])
        row += "  %s %6.1f" % (cfg, timeit(lambda: ops.conv3x3_p16_fwd(xp, wq, b, Cout, cfg, s)))
    print(row, flush=True)
