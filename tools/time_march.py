"""Dev tool: CUDA-event timing of the march / shade kernels at B=8 256x256 (not the bench)."""
import sys, os, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geomconsistentfr_b200 import ops
from geomconsistentfr_b200 import synthetic as O

B = int(sys.argv[1]) if len(sys.argv) > 1 else 8
depth = torch.zeros(B, 1, 256, 256); masks = torch.zeros(B, 256, 256, dtype=torch.uint8)
for b in range(B):
    d, m = O.synthetic_face(seed=b); depth[b, 0] = d; masks[b] = m
lights = torch.tensor([O.LIGHTS_18[i % 18] for i in range(B)], dtype=torch.float32)
P_L = O.light_point(lights)[1].cuda(); depth = depth.cuda(); bits = ops.mask_pack(masks.cuda())
albedo = torch.rand(B, 3, 256, 256, device="cuda"); amb = torch.full((B,), 0.3, device="cuda")
for variant in (1, 0):
    for _ in range(3): ops.shadow_march_fwd(depth, bits, P_L, variant=variant)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): dmin, _, _ = ops.shadow_march_fwd(depth, bits, P_L, variant=variant)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print("march variant %d: %.3f ms for B=%d  -> %.1f us/img, %.2f Gsamples/s, alg %.1f GB/s" % (
        variant, ms, B, 1e3 * ms / B, B * 160 * 65536 / ms / 1e6, B * 786432 / ms / 1e6))
for _ in range(3): ops.shade_render_fwd(albedo, depth, dmin, P_L, amb)
torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): ops.shade_render_fwd(albedo, depth, dmin, P_L, amb)
e1.record(); torch.cuda.synchronize()
print("shade_render: %.3f ms for B=%d" % (e0.elapsed_time(e1) / 20, B))
