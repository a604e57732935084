"""TEST INFRASTRUCTURE — generates tests/golden/* by running the UNMODIFIED reference
(/root/reference, via oracle/ref_shims.py) in the authoring container.  The GPU box has
no /root/reference, so the fixtures written here are committed.

    python -m oracle.make_golden            # from the repo root, ~3 min on 8 cores

Fixtures:
  model_epoch99.pth   the reference's shipped weights (model/model_epoch99.pth), byte copy
  ffhq.npz            10 FFHQ inputs (exactly the reference's cv2.resize(img/255) result, stored
                      as the integer 4-pixel sums q so that img = q/1020), the 10 skin masks, the
                      light used for each (TEST1:519-562), the 10 shipped result PNGs
                      (FFHQ_relighting_results/), and the reference forward's own outputs for 3 of them
                      (all ten outputs for 00508; shadow + rendered for 00295, 00695)
  march.npz           reference TEST1 forward with a synthetic depth injected at the depth head:
                      10 lights covering all 9 end-point branches (+5 bonus included) -> shadow weights
  train.npz           reference TRAIN forward (B=3, train-mode BN, predicted light) + autograd
                      gradients w.r.t. the depth map and the light/ambient head
"""
import os
import shutil
import sys

import cv2
import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shims, relight_oracle as O  # noqa: E402

REF = ref_shims.REFERENCE_ROOT
OUT = os.path.join(ROOT, "tests", "golden")

# image -> light (TEST1:519-562)
FFHQ_LIGHTS = {
    "00110": (0.7518, 0.0, 0.6594), "00300": (0.7518, 0.0, 0.6594), "00525": (0.7518, 0.0, 0.6594),
    "00104": (0.6893, 0.3991, 0.6047), "00290": (-0.5843, 0.0, 0.8115),
    "00322": (-0.7574, 0.0, 0.6529), "00572": (-0.7574, 0.0, 0.6529),
    "00695": (-0.5151, 0.4722, 0.7154), "00295": (0.0, 0.7071, 0.7071), "00508": (0.8138, -0.3420, 0.4698),
}
FULL_OUTPUT_FOR = ("00508", "00295", "00695")   # first: every output; others: shadow + rendered

MARCH_LIGHTS = [  # (tag, light) — all nine end-point branches of TEST1:363-437, and Lz = 0
    ("right_mid", (0.7518, 0.0, 0.6594)), ("right_above", (0.6893, 0.3991, 0.6047)),
    ("right_below", (0.8138, -0.3420, 0.4698)), ("left_mid", (-0.5843, 0.0, 0.8115)),
    ("left_above", (-0.5151, 0.4722, 0.7154)), ("left_below", (-0.8138, -0.3420, 0.4698)),
    ("mid_above", (0.0, 0.7071, 0.7071)), ("mid_below", (0.01, -0.7, 0.714)),
    ("inside", (0.01, 0.02, 0.9997)), ("grazing", (0.999, 0.047, 0.0)),
]


class _Fixed(nn.Module):
    """Stands in for a conv head so a chosen map can be injected into the reference forward."""

    def __init__(self, value):
        super().__init__()
        self.value = value

    def forward(self, x):
        return self.value


def _read_rgb(path):
    img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
    return img[:, :, ::-1].copy() if img.ndim == 3 else img


def load_ffhq():
    names = sorted(FFHQ_LIGHTS)
    q = np.zeros((len(names), 256, 256, 3), np.uint16)
    masks = np.zeros((len(names), 256, 256), np.uint8)
    pngs = np.zeros((len(names), 256, 256, 3), np.uint8)
    for i, n in enumerate(names):
        src = _read_rgb(os.path.join(REF, "sample_test_images_FFHQ", n + ".png"))
        ref_resized = cv2.resize(src / 255.0, (256, 256))                   # TEST1:515
        qi = np.rint(ref_resized * 1020.0)
        assert np.array_equal((qi / 1020.0).astype(np.float32), ref_resized.astype(np.float32))
        q[i] = qi.astype(np.uint16)
        masks[i] = cv2.imread(os.path.join(REF, "FFHQ_skin_masks", n + ".png"), cv2.IMREAD_UNCHANGED)
        pngs[i] = cv2.imread(os.path.join(REF, "FFHQ_relighting_results", n + "_rendered_image.png"))  # BGR as written
    lights = np.array([FFHQ_LIGHTS[n] for n in names], np.float32)
    return names, q, masks, lights, pngs


def main():
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    shutil.copyfile(os.path.join(REF, "model", "model_epoch99.pth"), os.path.join(OUT, "model_epoch99.pth"))
    K = O.intrinsic_matrix()

    # ---------------- ffhq.npz
    names, q, masks, lights, pngs = load_ffhq()
    net = ref_shims.reference_model("TEST1").eval()
    full = {}
    with torch.no_grad():
        for i, n in enumerate(names):
            if n not in FULL_OUTPUT_FOR:
                continue
            x = torch.from_numpy((q[i] / 1020.0)[None]).float()
            m = torch.from_numpy(masks[i].astype(np.float64).reshape(256, 256, 1)) / 255.0   # TEST1:586
            tl = torch.from_numpy(lights[i]).view(1, 3, 1, 1)
            r = net(x, 200, K, m, tl, torch.full((1, 1, 1), 0.5), m[None])
            keys = (("albedo", 0), ("depth", 1), ("shadow", 2), ("rendered", 5), ("ambient", 7), ("final", 8),
                    ("normals", 9)) if n == FULL_OUTPUT_FOR[0] else (("shadow", 2), ("rendered", 5), ("ambient", 7))
            for key, idx in keys:
                full["%s_%s" % (n, key)] = r[idx].numpy().astype(np.float32)
            print("ffhq", n, "done")
    np.savez_compressed(os.path.join(OUT, "ffhq.npz"), names=np.array(names), q=q, masks=masks, lights=lights,
                        pngs_bgr=pngs, **full)

    # ---------------- march.npz  (synthetic depth through the reference's own march)
    depth, mask01 = O.synthetic_face(seed=3)
    mask_u8 = (mask01.numpy() * 255).astype(np.uint8)
    # non-binary mask values like the shipped skin masks {0,64,128,255}: only "== 0" may matter
    mask_u8[100:120, 100:130] = np.where(mask_u8[100:120, 100:130] > 0, 64, 0)
    mask_u8[150:170, 110:150] = np.where(mask_u8[150:170, 110:150] > 0, 128, 0)
    mask_u8[60:70, 120:136] = 0                                             # a hole inside the face
    net = ref_shims.reference_model("TEST1").eval()
    net.conv_depth_c2_o = _Fixed((depth / 100.0).view(1, 1, 256, 256))
    m = torch.from_numpy(mask_u8.astype(np.float64).reshape(256, 256, 1)) / 255.0
    x = torch.zeros(1, 256, 256, 3)
    out = {"mask_u8": mask_u8}
    with torch.no_grad():
        for tag, L in MARCH_LIGHTS:
            tl = torch.tensor(L, dtype=torch.float32).view(1, 3, 1, 1)
            r = net(x, 200, K, m, tl, torch.full((1, 1, 1), 0.5), m[None])
            out["depth"] = r[1].numpy()[0, 0].astype(np.float32)          # the depth the march really saw
            out["shadow_" + tag] = r[2].numpy()[0].astype(np.float32)
            out["light_" + tag] = np.array(L, np.float32)
            print("march", tag, "done")
    out["tags"] = np.array([t for t, _ in MARCH_LIGHTS])
    np.savez_compressed(os.path.join(OUT, "march.npz"), **out)

    # ---------------- train.npz  (TRAIN forward + autograd gradients, B = 3 as hard-coded at TRAIN:358-359)
    torch.manual_seed(0)
    tnet = ref_shims.reference_model("TRAIN")      # batch_size 3, stays in train() mode like TRAIN:561-563
    sel = [names.index(n) for n in ("00110", "00290", "00295")]
    x = torch.from_numpy(q[sel] / 1020.0).float()
    mk = (masks[sel] > 128).astype(np.float64).reshape(3, 256, 256, 1)     # binarised like TRAIN:552-556, /255
    mt = torch.from_numpy(mk)
    r = tnet(x, 200, K, mt)
    albedo, depth_t, shadow, amb_l, full_s, rendered, unit_l, amb_v = r
    depth_t.retain_grad()
    m3 = mt.permute(0, 3, 1, 2).repeat(1, 3, 1, 1).float()
    target = x.permute(0, 3, 1, 2)
    comp = rendered * m3 + (1.0 - m3) * target
    loss = 20.0 * torch.sum((rendered * m3 - target * m3) ** 2) / torch.sum(m3) \
        + 8.0 * (1 - O.ssim(comp, target, data_range=1.0, size_average=True, nonnegative_ssim=True)) / 2.0   # TRAIN:633,643
    loss.backward()
    np.savez_compressed(
        os.path.join(OUT, "train.npz"), sel=np.array(sel), masks01=mk.astype(np.uint8)[..., 0],
        albedo=albedo.detach().numpy(), depth=depth_t.detach().numpy(), shadow=shadow.detach().numpy(),
        rendered=rendered.detach().numpy(), unit_light=unit_l.detach().numpy().reshape(3, 3),
        ambient=amb_v.detach().numpy().reshape(3), loss=np.float64(loss.item()),
        grad_depth=depth_t.grad.numpy(), grad_sl2_w=tnet.linear_SL2.weight.grad.numpy(),
        grad_sl2_b=tnet.linear_SL2.bias.grad.numpy(),
        grad_depth_head_w=tnet.conv_depth_c2_o.weight.grad.numpy(),
        grad_albedo_head_w=tnet.conv_albedo_c2_o.weight.grad.numpy())
    print("train done, loss", loss.item())
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))


if __name__ == "__main__":
    main()
