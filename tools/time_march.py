"""Times the three ray-march variants on the bench shapes (B = 8, 256x256, synthetic ellipse mask + FFHQ-like depth)."""
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
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for variant in (0, 1, 2):
    fn = lambda: ops.shadow_march_fwd(depth, bits, light, inside_bonus=5.0, variant=variant)
    for _ in range(3):
        fn()
    torch.cuda._sleep(20_000_000)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(20):
        fn()
    b.record()
    torch.cuda.synchronize()
    print("variant %d (%s): %.1f us per launch (20 back-to-back launches incl. the depth widening pass)" % (
        variant, ("thread-per-ray, culled + coarse skip", "thread-per-ray, literal", "warp-per-ray")[variant], 1e3 * a.elapsed_time(b) / 20))
