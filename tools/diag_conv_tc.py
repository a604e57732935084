"""Dev tool: error statistics of the tcgen05 conv vs fp64, per precision mode / NT / Cin."""
import sys, os, torch, torch.nn.functional as F
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geomconsistentfr_b200 import ops
torch.backends.cudnn.allow_tf32 = False
for (N, Cin, Cout, S) in [(2, 16, 16, 32), (2, 64, 64, 16), (2, 128, 64, 16), (2, 160, 64, 16), (2, 155, 155, 16)]:
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.randn(N, Cin, S, S, device="cuda", generator=g)
    w = torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / (3.0 * Cin ** 0.5)
    b = torch.randn(Cout, device="cuda", generator=g)
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=1)
    e32 = (F.conv2d(x, w, b, padding=1).double() - ref)
    print("Cin %d Cout %d: torch fp32 max %.2e mean %.2e" % (Cin, Cout, e32.abs().max(), e32.abs().mean()))
    for NT in (16, 32, 64):
        for prec in (1, 3):
            out = ops.c4_to_nchw(ops.conv3x3_tc_fwd(ops.nchw_to_c4(x), ops.conv_tc_pack_weights(w, NT), b, Cout, NT, act=None, precision=prec))
            if prec == 7 or prec == 3:
                e = out.double() - ref
            elif prec == 1:
                e = out.double() - ref   # single-pass
            else:
                e = out.double() - ref
            print("   NT %2d products %d: max %.2e mean %.2e signed-mean %.2e  corr(sign) %.3f" % (
                NT, prec, e.abs().max(), e.abs().mean(), e.mean(), (e * ref.sign()).mean() / e.abs().mean()))
