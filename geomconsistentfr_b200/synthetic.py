"""Synthetic workload generator for bench.py / tools: seeded 256x256 faces and the 18 Multi-PIE-style
light directions.  Pure host-side input generation; no reference code path and nothing from `oracle/` is involved
(tests/test_module_cpu.py checks that this generator and the oracle's agree, so both arms of the bench see the same
inputs)."""
import math

import torch

# 8 exact directions from the reference's source comments (TEST1:519-550) + 10 fitted from lighting_visualizations/
# (MP1..MP18 order, the fitted ones are labelled synthetic)
LIGHTS_18 = [
    (0.999, 0.047, 0.000), (0.965, 0.045, 0.259), (0.880, 0.041, 0.472), (0.7518, 0.0, 0.6594),
    (0.5145, 0.0, 0.8575), (0.285, 0.067, 0.956), (0.069, 0.000, 0.998), (-0.328, 0.066, 0.942),
    (-0.5843, 0.0, 0.8115), (-0.7574, 0.0, 0.6529), (-0.849, 0.046, 0.526), (-0.954, 0.051, 0.295),
    (-0.999, 0.054, 0.000), (0.6893, 0.3991, 0.6047), (0.4478, 0.4925, 0.7463), (0.000, 0.467, 0.884),
    (-0.5151, 0.4722, 0.7154), (-0.7076, 0.3892, 0.5897),
]


def synthetic_face(seed=0, H=256, W=256, noise=0.5):
    """Ellipsoidal face (semi-axes 0.3125 W x 0.390625 H, 120 deep) + a Gaussian nose bump (40) + seeded smooth
    pseudo-noise.  Returns depth (H,W) f32 and the face mask (H,W) u8 in {0,1}."""
    r = torch.arange(H, dtype=torch.float64).view(H, 1).expand(H, W)
    c = torch.arange(W, dtype=torch.float64).view(1, W).expand(H, W)
    rr = ((c - W / 2) / (W * 0.3125)) ** 2 + ((r - H / 2) / (H * 0.390625)) ** 2
    depth = 120.0 * torch.sqrt(torch.clamp(1.0 - rr, min=0.0)) \
        + 40.0 * torch.exp(-((c - W / 2) / 12.0) ** 2 - ((r - H * 0.546875) / 20.0) ** 2)
    ph = 0.61803398875 * (seed + 1)
    depth = depth + noise * (torch.sin(0.173 * c + 7.1 * ph) * torch.cos(0.211 * r + 3.3 * ph)
                             + 0.5 * torch.sin(0.057 * (c + 2 * r) + 1.7 * ph))
    return depth.float(), (rr < 1.0).to(torch.uint8)


def synthetic_batch(B, seed, H=256, W=256):
    """One bench batch: image U(0,1) NHWC f32 (seeded), one shared elliptical face mask u8 {0,255} [1,H,W], and B of the
    18 light directions [B,3,1,1]."""
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(B, H, W, 3, generator=g)
    _, m = synthetic_face(seed=seed, H=H, W=W)
    mask = (m * 255).to(torch.uint8).view(1, H, W)
    light = torch.tensor([LIGHTS_18[(seed + i) % 18] for i in range(B)], dtype=torch.float32).view(B, 3, 1, 1)
    return img, mask, light


def light_point(direction, distance=4013.0):
    """Point light at `distance` along the unit direction (TEST1:329-333).  direction: [B,3] tensor."""
    d = direction / direction.norm(dim=1, keepdim=True)
    return d, distance * d


assert len(LIGHTS_18) == 18 and all(abs(math.sqrt(sum(v * v for v in l)) - 1.0) < 2e-3 for l in LIGHTS_18)
