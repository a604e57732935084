"""One shape of the P16 conv, a few launches (for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from geomconsistentfr_b200 import ops
N, Cin, Cout, S = [int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (8, 16, 16, 256))]
cfg = tuple(int(v) for v in sys.argv[5:8]) if len(sys.argv) > 7 else ops.conv_p16_config(Cin, Cout, N, S, S)
x = torch.randn(N, Cin, S, S, device="cuda")
w = torch.randn(Cout, Cin, 3, 3, device="cuda") / (3 * Cin ** 0.5)
b = torch.zeros(Cout, device="cuda")
xp = ops.nchw_to_p16(x)
wq, s = ops.conv_p16_pack_weights(w, cfg[0], cfg[2])
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for _ in range(4):
    flush.zero_()
    ops.conv3x3_p16_fwd(xp, wq, b, Cout, cfg, s)
torch.cuda.synchronize()
print("done", cfg)
