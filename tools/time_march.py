"""Times the ray-march variants on the bench shapes (B = 8, 256x256, synthetic ellipse mask + FFHQ-like depth): the default
kernel in its four A/B configurations (warp shape x sample pairing, ops.march_config), the literal kernel and the warp-per-ray
mapping; stand-alone march (incl. the depth widening pass) and the fused march + shade launch the forward uses."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from geomconsistentfr_b200 import ops
from geomconsistentfr_b200.synthetic import LIGHTS_18, synthetic_face
B = 8
faces = [synthetic_face(seed=i, noise=2.0) for i in range(B)]
depth = torch.stack([d for d, _ in faces]).view(B, 1, 256, 256).cuda()
bits = ops.mask_pack((faces[0][1] * 255).view(1, 256, 256).cuda())
light = (4013.0 * torch.nn.functional.normalize(torch.tensor(LIGHTS_18[:B]), dim=1)).cuda()
albedo = torch.rand(B, 3, 256, 256, device="cuda")
amb = torch.full((B,), 0.4, device="cuda")


def timed(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda._sleep(20_000_000)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return 1e3 * a.elapsed_time(b) / n


for ws, ilp, order, cut in ((0, 1, 0, 0), (1, 1, 0, 0), (1, 2, 0, 0), (1, 2, 1, 0), (1, 2, 0, 1), (1, 2, 1, 1), (1, 3, 1, 1), (1, 4, 1, 1), (0, 2, 1, 1)):
    if True:
        ops.march_config(ws, ilp, order, cut)
        t_m = timed(lambda: ops.shadow_march_fwd(depth, bits, light, inside_bonus=5.0, variant=0))
        t_f = timed(lambda: ops.march_shade_fwd(albedo, depth, bits, light, amb, inside_bonus=5.0))
        print("variant 0, warp %s, samples %s, %s: march %.1f us, fused march+shade %.1f us (20 back-to-back launches incl. the depth widening pass)"
              % (("32x1", "8x4")[ws], ("one by one", "in pairs", "in threes", "in fours")[ilp - 1], ("tile-major order", "light-aware order")[order] + (", early cut-off" if cut else ""), t_m, t_f))
ops.march_config(-1, 0, -1, -1)
for variant in (1, 2):
    t = timed(lambda: ops.shadow_march_fwd(depth, bits, light, inside_bonus=5.0, variant=variant))
    print("variant %d (%s): %.1f us per launch" % (variant, ("", "thread-per-ray, literal", "warp-per-ray")[variant], t))
