"""TEST INFRASTRUCTURE — the CPU oracle.  Not product code.

A CPU (torch, fp32/fp64 exactly where the reference is fp32/fp64) restatement of the
hot path of andrewhou1/GeomConsistentFR: the hard-shadow ray-march, the Lambertian
shading/render and the RelightNet / PatchGAN CNNs.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import it, and only as the checker or the timed CPU baseline; the product package
(`geomconsistentfr_b200/`) never does.

Citations are relative to /root/reference:
  TRAIN = train_raytracing_relighting_CelebAHQ_DSSIM_8x.py
  TEST1 = test_relight_single_image.py

Pinning: oracle/make_golden.py runs the UNMODIFIED reference through
oracle/ref_shims.py in the authoring container and commits its outputs under
tests/golden/; tests/test_oracle_golden.py checks this restatement against them, and
against the 10 relit PNGs the reference ships (FFHQ_relighting_results/, mask interior,
<= 1/255).  Two upstream dependencies are absent from /root/reference and are restated
from their published algorithms: kornia 0.4.1 `depth_to_normals` (pinned indirectly by
the 10 PNGs) and pytorch_msssim `ssim` (version pinned nowhere by the reference ->
DSSIM parity is UNPINNED; the restatement is the oracle).
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

IMG = 256
LIGHT_DISTANCE = 4013.0          # TRAIN:47
DIRECTIONAL_INTENSITY = 0.5      # TRAIN:46
NUM_SAMPLES = 160                # TRAIN:48
T0, DT = 0.025, 0.005            # TRAIN:468  np.arange(0.025, 0.825, 0.005)
FOCAL, DEPTH_OFFSET = 1570.0, 1610.0   # TRAIN:353,572-573
# lighting-transfer variant (TEST_LT = test_relight_single_image_lighting_transfer.py): TEST_LT:20,22,325,451,530-531,332
LT = dict(directional_intensity=0.41, num_samples=159, t0=0.03, depth_offset=1410.0, focal=700.0, z_floor=0.16)


# --------------------------------------------------------------------------------------
# un-vendored dependencies, restated from upstream
# --------------------------------------------------------------------------------------
def depth_to_normals(depth, camera_matrix, normalize_points=False):
    """kornia 0.4.1 `kornia.geometry.depth.depth_to_normals` (call sites TRAIN:353, TEST1:326).

    depth (B,1,H,W); camera_matrix (1|B,3,3).  xyz = ((u-cx)/fx, (v-cy)/fy, 1) * depth with
    u = column, v = row; spatial gradient = normalised Sobel (kernel / 8), replicate padding,
    cross-correlation; normals = normalize(cross(d/du, d/dv))."""
    B, _, H, W = depth.shape
    K = camera_matrix.to(depth.dtype)
    dev = depth.device
    ys, xs = torch.meshgrid(torch.arange(H, dtype=depth.dtype, device=dev), torch.arange(W, dtype=depth.dtype, device=dev),
                            indexing="ij")
    fx, fy = K[:, 0, 0].view(-1, 1, 1), K[:, 1, 1].view(-1, 1, 1)
    cx, cy = K[:, 0, 2].view(-1, 1, 1), K[:, 1, 2].view(-1, 1, 1)
    x = (xs[None] - cx) / fx
    y = (ys[None] - cy) / fy
    ones = torch.ones_like(x)
    pts = torch.stack([x, y, ones], 1).to(depth.dtype).expand(B, 3, H, W)
    if normalize_points:
        pts = F.normalize(pts, dim=1, p=2)
    xyz = pts * depth
    kx = torch.tensor([[-1.0, 0.0, 1.0], [-2.0, 0.0, 2.0], [-1.0, 0.0, 1.0]], dtype=depth.dtype, device=dev) / 8.0
    k = torch.stack([kx, kx.t()])[:, None]
    g = F.conv2d(F.pad(xyz.reshape(B * 3, 1, H, W), (1, 1, 1, 1), mode="replicate"), k).view(B, 3, 2, H, W)
    a, b = g[:, :, 0], g[:, :, 1]
    return F.normalize(torch.cross(a, b, dim=1), dim=1, p=2)


def ssim(X, Y, data_range=255, size_average=True, win_size=11, win_sigma=1.5, win=None,
         K=(0.01, 0.03), nonnegative_ssim=False):
    """pytorch_msssim.ssim (call site TRAIN:643: data_range=1.0, size_average=True,
    nonnegative_ssim=True).  Separable VALID 11-tap Gaussian per channel."""
    c = torch.arange(win_size, dtype=torch.float32) - win_size // 2
    w = torch.exp(-(c ** 2) / (2 * win_sigma ** 2))
    w = (w / w.sum()).to(X.dtype).to(X.device)

    def gf(x):
        C = x.shape[1]
        x = F.conv2d(x, w.view(1, 1, -1, 1).repeat(C, 1, 1, 1), groups=C)
        return F.conv2d(x, w.view(1, 1, 1, -1).repeat(C, 1, 1, 1), groups=C)

    C1, C2 = (K[0] * data_range) ** 2, (K[1] * data_range) ** 2
    mu1, mu2 = gf(X), gf(Y)
    s1, s2, s12 = gf(X * X) - mu1 * mu1, gf(Y * Y) - mu2 * mu2, gf(X * Y) - mu1 * mu2
    cs = (2 * s12 + C2) / (s1 + s2 + C2)
    per = (((2 * mu1 * mu2 + C1) / (mu1 * mu1 + mu2 * mu2 + C1)) * cs).flatten(2).mean(-1)
    if nonnegative_ssim:
        per = torch.relu(per)
    return per.mean() if size_average else per.mean(1)


# --------------------------------------------------------------------------------------
# geometry: pixel grid, light point, ray end points
# --------------------------------------------------------------------------------------
def pixel_grid(H=IMG, W=IMG, device="cpu"):
    """TRAIN:51-55: xx = col - W/2, yy = H/2 - row (fp32)."""
    cols = torch.arange(W, dtype=torch.float32, device=device)[None, :].expand(H, W)
    rows = torch.arange(H, dtype=torch.float32, device=device)[:, None].expand(H, W)
    return (cols - W / 2.0).contiguous(), (H / 2.0 - rows).contiguous()


def light_point(light_dir, clamp_z=False, light_distance=LIGHT_DISTANCE):
    """light_dir (B,3) -> (unit (B,3), point (B,3)).  TEST1:332-335; train clamps z>=0 first (TRAIN:357-362)."""
    L = light_dir.reshape(-1, 3)
    if clamp_z:
        L = torch.cat((L[:, 0:2], torch.clamp(L[:, 2:3], min=0.0)), 1)
    unit = F.normalize(L, p=2, dim=1)
    return unit, light_distance * unit


def ray_endpoints(xx, yy, Lx, Ly):
    """TRAIN:375-465 / TEST1:352-442.  xx,yy (H,W) fp32 ; Lx,Ly 0-d fp32 tensors.
    Returns end points (2,H,W) fp32, clamped to x in [-128,127], y in [-127,128]."""
    H, W = xx.shape
    xmin, xmax = -(W / 2.0), W - W / 2.0 - 1
    ymin, ymax = 1 - H / 2.0, H / 2.0
    slopes = (Ly - yy) / (Lx - xx + 0.0001)
    intercepts = Ly - slopes * Lx
    lx, ly = float(Lx.detach()) if torch.is_tensor(Lx) else float(Lx), float(Ly.detach()) if torch.is_tensor(Ly) else float(Ly)

    def x_edge(xe):
        x = torch.full((H, W), xe, dtype=torch.float32, device=xx.device)
        return x, slopes * x + intercepts

    def y_edge(ye):
        y = torch.full((H, W), ye, dtype=torch.float32, device=xx.device)
        return (y - intercepts) / (slopes + 0.0001), y

    xe = xmin if lx < xmin else (xmax if lx > xmax else None)
    ye = ymin if ly < ymin else (ymax if ly > ymax else None)
    if xe is not None and ye is not None:
        ex_x, ex_y = x_edge(xe)
        ey_x, ey_y = y_edge(ye)
        hit_y = torch.logical_and(ey_x >= xmin, ey_x <= xmax)
        ex = torch.where(hit_y, ey_x, ex_x)
        ey = torch.where(hit_y, ey_y, ex_y)
    elif xe is not None:
        ex, ey = x_edge(xe)
    elif ye is not None:
        ex, ey = y_edge(ye)
    else:
        ex = torch.full((H, W), lx, dtype=torch.float32, device=xx.device)
        ey = torch.full((H, W), ly, dtype=torch.float32, device=xx.device)
    ex = torch.clamp(ex, xmin, xmax)         # TRAIN:462-465 (hard-coded -128/127/-127/128 for 256x256)
    ey = torch.clamp(ey, ymin, ymax)
    return torch.stack((ex, ey), 0)


def light_inside_image(Lx, Ly, H=IMG, W=IMG, rule="image"):
    """TEST1:495 (rule "image": the light projects inside the image rectangle) / TEST_LT:503 (rule "wide": inside
    +-4 image sizes)."""
    lx, ly = float(Lx.detach()) if torch.is_tensor(Lx) else float(Lx), float(Ly.detach()) if torch.is_tensor(Ly) else float(Ly)
    if rule == "wide":
        return (-4 * W <= lx <= 4 * W) and (4 * (1 - H) <= ly <= 4 * H)
    return (-(W / 2.0) <= lx <= W - W / 2.0 - 1) and (1 - H / 2.0 <= ly <= H / 2.0)


def sample_increments(t0=T0, dt=DT, n=NUM_SAMPLES):
    """TRAIN:468 / TEST_LT:451: exactly np.arange's fp64 values start + k*step."""
    if (t0, dt, n) == (T0, DT, NUM_SAMPLES):
        t = np.arange(0.025, 0.825, 0.005)
        assert t.shape[0] == NUM_SAMPLES
        return torch.from_numpy(t)
    if (t0, dt, n) == (LT["t0"], DT, LT["num_samples"]):
        t = np.arange(0.03, 0.825, 0.005)
        assert t.shape[0] == LT["num_samples"]
        return torch.from_numpy(t)
    return torch.from_numpy(t0 + dt * np.arange(n, dtype=np.float64))


# --------------------------------------------------------------------------------------
# the ray-march: minimum point-to-ray distance over the samples
# --------------------------------------------------------------------------------------
def _march_one(depth, mask2d, P_L, xx, yy, t, inside_bonus, chunk, inside_rule="image"):
    """One image.  depth (H,W) f32; mask2d (H,W) any dtype (tested ==0); P_L (3,) f32."""
    H, W = depth.shape
    end = ray_endpoints(xx, yy, P_L[0], P_L[1])
    start = torch.stack((xx, yy), 0)                       # (2,H,W) f32
    diff = end - start                                     # TRAIN:467 f32
    B3 = torch.stack((xx, yy, depth), 0)                   # points_B (3,H,W) f32
    BC = P_L.view(3, 1, 1) - B3                            # TRAIN:505-507
    den = torch.sqrt(torch.sum(BC * BC, dim=0) + 0.0001)   # TRAIN:509
    mins, args = [], []
    for k0 in range(0, t.shape[0], chunk):
        tk = t[k0:k0 + chunk].view(-1, 1, 1, 1)            # fp64
        p = start[None] + tk * diff[None]                  # fp64 (n,2,H,W) TRAIN:472,480
        # nearest index for the face-mask test, TRAIN:472-477 (round half to even, twice)
        pr = torch.round(p)
        ci = torch.round(pr[:, 0] + (W / 2.0)).long()
        ri = torch.round((H / 2.0) - pr[:, 1]).long()
        # continuous index, TRAIN:480-487
        u = (p[:, 0] + (W / 2.0)) - 0.0001
        v = ((H / 2.0) - p[:, 1]) - 0.0001
        uc, uf = torch.ceil(u).int(), torch.floor(u).int()
        vc, vf = torch.ceil(v).int(), torch.floor(v).int()
        ufl, ucl, vfl, vcl = uf.long(), uc.long(), vf.long(), vc.long()   # negative index wraps (python)
        d_ul, d_ur = depth[vfl, ufl], depth[vfl, ucl]      # TRAIN:488-491
        d_ll, d_lr = depth[vcl, ufl], depth[vcl, ucl]
        up = d_ul * (uc - u) + d_ur * (u - uf)             # fp64, TRAIN:492
        lo = d_ll * (uc - u) + d_lr * (u - uf)             # TRAIN:493
        zi = up * (vc - v) + lo * (v - vf)                 # TRAIN:494
        A = torch.stack((u - (W / 2.0), (H / 2.0) - v, zi), 0).float()   # TRAIN:497-502
        BA = A - B3[:, None]
        cr = torch.cross(BA, BC[:, None].expand_as(BA), dim=0)          # TRAIN:508
        dist = torch.sqrt(torch.sum(cr * cr, dim=0) + 0.0001) / den[None]
        outside = mask2d[ri, ci] == 0                      # TRAIN:510
        dist = torch.logical_not(outside) * dist + outside * 1000000.0  # TRAIN:512
        vmin, imin = torch.min(dist, dim=0)                # TRAIN:514
        mins.append(vmin)
        args.append(imin + k0)
    allmin = torch.stack(mins, 0)
    dmin, which = torch.min(allmin, dim=0)
    arg = torch.gather(torch.stack(args, 0), 0, which[None])[0]
    if inside_bonus != 0.0 and light_inside_image(P_L[0], P_L[1], H, W, inside_rule):
        dmin = dmin + inside_bonus                         # TEST1:495-496
    return dmin, arg


def shadow_march(depth, mask, light_pt, t0=T0, dt=DT, n=NUM_SAMPLES, inside_bonus=0.0, chunk=32,
                 return_argmin=False, inside_rule="image"):
    """depth (B,1,H,W) f32; mask (B|1,H,W) (any dtype; ==0 means outside the face);
    light_pt (B,3) f32 = 4013*unit(L).  Returns d_min (B,H,W) f32 (TRAIN:374-515)."""
    B, _, H, W = depth.shape
    xx, yy = pixel_grid(H, W, depth.device)
    t = sample_increments(t0, dt, n).to(depth.device)
    outs, args = [], []
    for i in range(B):
        m = mask[i if mask.shape[0] > 1 else 0]
        d, a = _march_one(depth[i, 0], m, light_pt[i], xx, yy, t, inside_bonus, chunk, inside_rule)
        outs.append(d)
        args.append(a)
    d = torch.stack(outs, 0)
    return (d, torch.stack(args, 0)) if return_argmin else d


def shadow_weight(d_min):
    """TRAIN:517 — 1 - 4e^-d/(1+e^-d)^2  (== tanh^2(d/2))."""
    return -4 * torch.exp(-d_min) / torch.pow((1 + torch.exp(-d_min)), 2) + 1


# --------------------------------------------------------------------------------------
# normals, Lambertian shading, render
# --------------------------------------------------------------------------------------
def intrinsic_matrix(H=IMG, W=IMG, focal=FOCAL):
    """TRAIN:571-577 (float64, shape (1,3,3))."""
    K = np.zeros((1, 3, 3))
    K[:, 0, 0] = focal
    K[:, 1, 1] = focal
    K[:, 2, 2] = 1.0
    K[:, 0, 2] = W / 2.0
    K[:, 1, 2] = H / 2.0
    return torch.from_numpy(K)


def shade(depth, K, light_pt, ambient_values, depth_offset=DEPTH_OFFSET, intensity=DIRECTIONAL_INTENSITY):
    """TRAIN:353-369.  depth (B,1,H,W); light_pt (B,3); ambient_values (B,).
    Returns normals (B,3,H,W) [after the y flip and 2nd normalise], directional (B,H,W),
    ambient_light (B,H,W), full_shading (B,H,W)."""
    B, _, H, W = depth.shape
    xx, yy = pixel_grid(H, W, depth.device)
    n = depth_to_normals(depth + depth_offset, K.to(depth.device))
    n = torch.cat((n[:, 0:1], -n[:, 1:2], n[:, 2:3]), 1)                  # TRAIN:354
    P = torch.cat((xx.expand(B, 1, H, W), yy.expand(B, 1, H, W), depth), 1)
    l = F.normalize(light_pt.view(B, 3, 1, 1) - P, p=2, dim=1)            # TRAIN:364
    n = F.normalize(n, p=2, dim=1)                                        # TRAIN:365
    directional = intensity * torch.clamp(torch.sum(n * l, dim=1), min=0.0)
    ambient_light = ambient_values.view(B, 1, 1).expand(B, H, W)
    return n, directional, ambient_light, ambient_light + directional


def render(albedo, shadow, full_shading, ambient_light):
    """TRAIN:518-522."""
    final = shadow * full_shading + (1 - shadow) * ambient_light
    return final, albedo * final[:, None]


def relight_from_maps(albedo, depth, mask, light_dir, ambient_values, K=None, inside_bonus=5.0,
                      clamp_z=False, t0=T0, dt=DT, n=NUM_SAMPLES, depth_offset=DEPTH_OFFSET,
                      intensity=DIRECTIONAL_INTENSITY, inside_rule="image"):
    """Everything after the CNN (TEST1:325-505): maps -> dict of outputs."""
    if K is None:
        K = intrinsic_matrix(depth.shape[2], depth.shape[3])
    unit, P_L = light_point(light_dir, clamp_z)
    normals, directional, amb, full = shade(depth, K, P_L, ambient_values, depth_offset, intensity)
    d_min = shadow_march(depth, mask, P_L, t0, dt, n, inside_bonus, inside_rule=inside_rule)
    s = shadow_weight(d_min)
    final, rendered = render(albedo, s, full, amb)
    return dict(d_min=d_min, shadow=s, ambient_light=amb, full_shading=full, final_shading=final,
                rendered=rendered, normals=normals, unit_light=unit.view(-1, 3, 1, 1))


# --------------------------------------------------------------------------------------
# the CNNs (plain torch fp32) — same parameter names as the reference (TRAIN:58-194)
# --------------------------------------------------------------------------------------
ENCODER_LAYERS = [  # (name, cin, cout, k)   TRAIN:58-70
    ("conv_c1_og", 3, 16, 5), ("conv_h1_1", 16, 16, 3), ("conv_h1_2", 16, 16, 3),
    ("conv_h2_1", 16, 32, 3), ("conv_h2_2", 32, 32, 3), ("conv_shortcut_h1_out", 16, 32, 3),
    ("conv_h3_1", 32, 64, 3), ("conv_h3_2", 64, 64, 3), ("conv_shortcut_h2_out", 32, 64, 3),
    ("conv_h4_1", 64, 155, 3), ("conv_h4_2", 155, 155, 3), ("conv_shortcut_h3_out", 64, 155, 3),
]


def decoder_layers(p):  # (name, kind, cin, cout, k, has_bn)   TRAIN:91-114
    return [
        ("deconv_%s_h5_1" % p, "deconv", 128, 64, 3, True), ("deconv_%s_h5_2" % p, "deconv", 64, 64, 3, True),
        ("deconv_%s_shortcut_all_features" % p, "deconv", 128, 64, 3, True),
        ("conv_%s_skip_s1_1" % p, "conv", 64, 64, 3, True), ("conv_%s_skip_s1_2" % p, "conv", 64, 64, 3, True),
        ("deconv_%s_h6_1" % p, "deconv", 64, 32, 3, True), ("deconv_%s_h6_2" % p, "deconv", 32, 32, 3, True),
        ("deconv_%s_shortcut_h5_out" % p, "deconv", 64, 32, 3, True),
        ("conv_%s_skip_s2_1" % p, "conv", 32, 32, 3, True), ("conv_%s_skip_s2_2" % p, "conv", 32, 32, 3, True),
        ("deconv_%s_h7_1" % p, "deconv", 32, 16, 3, True), ("deconv_%s_h7_2" % p, "deconv", 16, 16, 3, True),
        ("deconv_%s_shortcut_h6_out" % p, "deconv", 32, 16, 3, True),
        ("conv_%s_skip_s3_1" % p, "conv", 16, 16, 3, True), ("conv_%s_skip_s3_2" % p, "conv", 16, 16, 3, True),
        ("deconv_%s_h8_1" % p, "deconv", 16, 16, 3, True), ("deconv_%s_h8_2" % p, "deconv", 16, 16, 3, True),
        ("conv_%s_skip_s4_1" % p, "conv", 16, 16, 3, True), ("conv_%s_skip_s4_2" % p, "conv", 16, 16, 3, True),
        ("conv_%s_c2_1" % p, "conv", 16, 16, 3, True), ("conv_%s_c2_2" % p, "conv", 16, 16, 1, True),
        ("conv_%s_c2_3" % p, "conv", 16, 16, 1, True),
        ("conv_%s_c2_o" % p, "conv", 16, 3 if p == "albedo" else 1, 1, False),
    ]


def _bn_name(layer):
    return "bn_" + layer.split("_", 1)[1]


class RelightNetOracle(nn.Module):
    """RelightNet (TRAIN:38-350 / TEST1:12-323) in plain torch; strict state_dict compatibility
    with model/model_epoch99.pth."""

    def __init__(self, variant="default"):
        """variant "lighting_transfer": the nine shortcut (de)convs are 1x1 and bias-free (TEST_LT:36-42,66-76,119-129;
        strict state_dict compatibility with model_lighting_transfer/model_epoch106.pth, 391 tensors)."""
        super().__init__()
        assert variant in ("default", "lighting_transfer")
        self.variant = variant
        lt = variant == "lighting_transfer"

        def make(mod, name, cin, cout, k):
            if lt and "shortcut" in name:
                return mod(cin, cout, 1, bias=False)
            return mod(cin, cout, k, padding=(k // 2, k // 2))

        for name, cin, cout, k in ENCODER_LAYERS:
            setattr(self, name, make(nn.Conv2d, name, cin, cout, k))
            setattr(self, _bn_name(name), nn.BatchNorm2d(cout))
        self.linear_SL1 = nn.Linear(27, 128)
        self.linear_SL2 = nn.Linear(128, 4)
        for p in ("albedo", "depth"):
            for name, kind, cin, cout, k, has_bn in decoder_layers(p):
                setattr(self, name, make(nn.ConvTranspose2d if kind == "deconv" else nn.Conv2d, name, cin, cout, k))
                if has_bn:
                    setattr(self, _bn_name(name), nn.BatchNorm2d(cout))

    def _cb(self, name, x):
        return getattr(self, _bn_name(name))(getattr(self, name)(x))

    def _decoder(self, p, idf, skips, epoch):
        c1_og, h1_og, h2_og, h3_og = skips
        L = lambda x: F.leaky_relu(x, 0.2)
        up = lambda x: F.interpolate(x, scale_factor=2, mode="nearest")
        cb = self._cb
        h = up(L(cb("deconv_%s_shortcut_all_features" % p, idf) +
                 cb("deconv_%s_h5_2" % p, L(cb("deconv_%s_h5_1" % p, idf)))))
        s = L(h3_og + cb("conv_%s_skip_s1_2" % p, L(cb("conv_%s_skip_s1_1" % p, h3_og))))
        if epoch > 8:
            h = h + s
        h = up(L(cb("deconv_%s_shortcut_h5_out" % p, h) +
                 cb("deconv_%s_h6_2" % p, L(cb("deconv_%s_h6_1" % p, h)))))
        s = L(h2_og + cb("conv_%s_skip_s2_2" % p, L(cb("conv_%s_skip_s2_1" % p, h2_og))))
        if epoch > 10:
            h = h + s
        h = up(L(cb("deconv_%s_shortcut_h6_out" % p, h) +
                 cb("deconv_%s_h7_2" % p, L(cb("deconv_%s_h7_1" % p, h)))))
        s = L(h1_og + cb("conv_%s_skip_s3_2" % p, L(cb("conv_%s_skip_s3_1" % p, h1_og))))
        if epoch > 12:
            h = h + s
        h = up(L(h + cb("deconv_%s_h8_2" % p, L(cb("deconv_%s_h8_1" % p, h)))))
        s = L(c1_og + cb("conv_%s_skip_s4_2" % p, L(cb("conv_%s_skip_s4_1" % p, c1_og))))
        if epoch > 14:
            h = h + s
        h = L(cb("conv_%s_c2_1" % p, h))
        h = L(cb("conv_%s_c2_2" % p, h))
        h = L(cb("conv_%s_c2_3" % p, h))
        return getattr(self, "conv_%s_c2_o" % p)(h)

    def cnn(self, img_nhwc, epoch):
        """-> albedo (B,3,H,W) after sigmoid, depth (B,1,H,W) after x100, SL_lin2 (B,4)."""
        L = lambda x: F.leaky_relu(x, 0.2)
        cb = self._cb
        x = img_nhwc.permute(0, 3, 1, 2)
        c1_og = L(cb("conv_c1_og", x))
        c1 = F.max_pool2d(c1_og, 2)
        h1_og = L(c1 + cb("conv_h1_2", L(cb("conv_h1_1", c1))))
        h1 = F.max_pool2d(h1_og, 2)
        h2_og = L(cb("conv_shortcut_h1_out", h1) + cb("conv_h2_2", L(cb("conv_h2_1", h1))))
        h2 = F.max_pool2d(h2_og, 2)
        h3_og = L(cb("conv_shortcut_h2_out", h2) + cb("conv_h3_2", L(cb("conv_h3_1", h2))))
        h3 = F.max_pool2d(h3_og, 2)
        h4 = L(cb("conv_shortcut_h3_out", h3) + cb("conv_h4_2", L(cb("conv_h4_1", h3))))
        idf, lf = h4[:, 0:128], h4[:, 128:155]
        pooled = F.avg_pool2d(lf, (16, 16), (1, 1)).flatten(1)                # TRAIN:85,230
        sl = self.linear_SL2(L(self.linear_SL1(pooled)))                   # (B,4)
        skips = (c1_og, h1_og, h2_og, h3_og)
        albedo = torch.sigmoid(self._decoder("albedo", idf, skips, epoch))
        depth = 100.0 * self._decoder("depth", idf, skips, epoch)
        return albedo, depth, sl

    def forward_test(self, img, epoch, K, mask, target_lighting, ambient_offset=-0.1, inside_bonus=5.0):
        """TEST1:169-505 semantics.  mask (H,W,1) shared by the batch.  Returns the 10-tuple."""
        albedo, depth, sl = self.cnn(img, epoch)
        B = img.shape[0]
        amb = sl[:, 0] + ambient_offset                                    # TEST1:342
        m = mask.reshape(1, mask.shape[0], mask.shape[1])
        o = relight_from_maps(albedo, depth, m, target_lighting.reshape(B, 3), amb, K, inside_bonus)
        return (albedo, depth, o["shadow"], o["ambient_light"], o["full_shading"], o["rendered"],
                o["unit_light"], amb.view(B, 1, 1), o["final_shading"], o["normals"])

    def forward_lighting_transfer(self, img, epoch, K, mask, target_lighting, target_ambient_values):
        """TEST_LT:169-514 semantics (variant "lighting_transfer").  mask (H,W,1); target_lighting (B,3,1,1);
        target_ambient_values (B,1,1) — USED here (TEST_LT:348), unlike TEST1.  Returns the 12-tuple: TEST1's ten +
        estimated_unit_light_direction (B,3,1,1) [z floored at 0.16, TEST_LT:332-334] + estimated_ambient_light (B,1,1)."""
        albedo, depth, sl = self.cnn(img, epoch)
        B = img.shape[0]
        est = torch.cat((sl[:, 1:3], torch.maximum(sl[:, 3:4], torch.tensor(LT["z_floor"]))), 1)
        est_unit = F.normalize(est, p=2, dim=1).view(B, 3, 1, 1)
        amb = target_ambient_values.reshape(B).float()
        m = mask.reshape(1, mask.shape[0], mask.shape[1])
        o = relight_from_maps(albedo, depth, m, target_lighting.reshape(B, 3), amb, K, 5.0, t0=LT["t0"], n=LT["num_samples"],
                              depth_offset=LT["depth_offset"], intensity=LT["directional_intensity"], inside_rule="wide")
        return (albedo, depth, o["shadow"], o["ambient_light"], o["full_shading"], o["rendered"], o["unit_light"],
                amb.view(B, 1, 1), o["final_shading"], o["normals"], est_unit, sl[:, 0].view(B, 1, 1))

    def forward_train(self, img, epoch, K, masks):
        """TRAIN:196-524 semantics.  masks (B,H,W,1).  Returns the 8-tuple."""
        albedo, depth, sl = self.cnn(img, epoch)
        B = img.shape[0]
        amb = sl[:, 0]
        m = masks.reshape(B, masks.shape[1], masks.shape[2])
        o = relight_from_maps(albedo, depth, m, sl[:, 1:4], amb, K, 0.0, clamp_z=True)
        return (albedo, depth, o["shadow"], o["ambient_light"], o["full_shading"], o["rendered"],
                o["unit_light"], amb.view(B, 1, 1))


class PatchGANOracle(nn.Module):
    """TRAIN:15-35."""

    def __init__(self):
        super().__init__()
        self.conv1 = nn.Conv2d(3, 64, 4, stride=2, padding=(1, 1))
        self.conv2 = nn.Conv2d(64, 128, 4, stride=2, padding=(1, 1))
        self.conv3 = nn.Conv2d(128, 256, 4, stride=2, padding=(1, 1))
        self.conv4 = nn.Conv2d(256, 512, 4, stride=2, padding=(1, 1))
        self.conv5 = nn.Conv2d(512, 1, 4, stride=1, padding=(1, 1))
        self.bn2, self.bn3, self.bn4 = nn.BatchNorm2d(128), nn.BatchNorm2d(256), nn.BatchNorm2d(512)

    def forward(self, img):
        x = F.leaky_relu(self.conv1(img), 0.2)
        x = F.leaky_relu(self.bn2(self.conv2(x)), 0.2)
        x = F.leaky_relu(self.bn3(self.conv3(x)), 0.2)
        x = F.leaky_relu(self.bn4(self.conv4(x)), 0.2)
        return self.conv5(x)


# --------------------------------------------------------------------------------------
# synthetic inputs — deterministic, no RNG-version dependence
# --------------------------------------------------------------------------------------
def synthetic_face(seed=0, H=IMG, W=IMG, noise=0.5):
    """Ellipsoidal face + nose bump + smooth pseudo-noise.  Returns depth (H,W) f32, mask (H,W) u8."""
    r, c = np.meshgrid(np.arange(H, dtype=np.float64), np.arange(W, dtype=np.float64), indexing="ij")
    rr = ((c - W / 2) / (W * 0.3125)) ** 2 + ((r - H / 2) / (H * 0.390625)) ** 2
    depth = 120.0 * np.sqrt(np.maximum(1.0 - rr, 0.0)) + 40.0 * np.exp(-((c - W / 2) / 12.0) ** 2 - ((r - H * 0.546875) / 20.0) ** 2)
    # smooth deterministic "noise": a few incommensurate sinusoids keyed by the seed
    ph = 0.61803398875 * (seed + 1)
    depth = depth + noise * (np.sin(0.173 * c + 7.1 * ph) * np.cos(0.211 * r + 3.3 * ph) + 0.5 * np.sin(0.057 * (c + 2 * r) + 1.7 * ph))
    mask = (rr < 1.0).astype(np.uint8)
    return torch.from_numpy(depth.astype(np.float32)), torch.from_numpy(mask)


# the 18 Multi-PIE-style light directions (8 exact from TEST1:519-550, 10 fitted)
LIGHTS_18 = [
    (0.999, 0.047, 0.000), (0.965, 0.045, 0.259), (0.880, 0.041, 0.472), (0.7518, 0.0, 0.6594),
    (0.5145, 0.0, 0.8575), (0.285, 0.067, 0.956), (0.069, 0.000, 0.998), (-0.328, 0.066, 0.942),
    (-0.5843, 0.0, 0.8115), (-0.7574, 0.0, 0.6529), (-0.849, 0.046, 0.526), (-0.954, 0.051, 0.295),
    (-0.999, 0.054, 0.000), (0.6893, 0.3991, 0.6047), (0.4478, 0.4925, 0.7463), (0.000, 0.467, 0.884),
    (-0.5151, 0.4722, 0.7154), (-0.7076, 0.3892, 0.5897),
]
