"""TEST INFRASTRUCTURE — generates tests/golden/planes.npz by EXECUTING the reference's own export block, the source lines
test_raytracing_relighting_CelebAHQ_DSSIM_8x.py:584-608 (TESTB: permutes, depth normalisation, the mask products and the six
cv2.imwrite calls), unmodified, on synthetic forward outputs.  `cv2.imwrite` is intercepted and the array it is handed goes
through cv2's real PNG encoder and back, so the fixture holds exactly the bytes the reference would have written.  It pins
oracle/postprocess_oracle.py (export_planes_u8, composite_bgr_u8 and their u8 rounding).

    python -m oracle.make_golden_planes          # from the repo root, < 1 s
"""
import os
import sys
import textwrap
import types

import cv2
import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shims  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
SCRIPT = os.path.join(ref_shims.REFERENCE_ROOT, "test_raytracing_relighting_CelebAHQ_DSSIM_8x.py")
FIRST, LAST = 584, 608


def main():
    H = W = 64
    g = np.random.default_rng(11)
    inputs = dict(
        albedo=g.random((1, 3, H, W)).astype(np.float32), depth=(g.standard_normal((1, 1, H, W)) * 60 + 40).astype(np.float32),
        shadow=g.random((1, H, W)).astype(np.float32), final=(g.random((1, H, W)) * 1.2).astype(np.float32),
        normals=(g.random((1, 3, H, W)) * 2 - 1).astype(np.float32),
        rendered=(g.random((1, 3, H, W)) * 1.3 - 0.1).astype(np.float32),          # out of range on both sides: saturation
        image=g.random((1, H, W, 3)), mask_u8=g.choice(np.array([0, 64, 128, 255], np.uint8), size=(H, W)))
    inputs["rendered"][0, :, 3, :8] = (np.arange(8, dtype=np.float32) + 0.5) / 255.0        # .5 ties after * 255
    lines = open(SCRIPT).read().split("\n")[FIRST - 1:LAST]
    block = textwrap.dedent("\n".join(lines))
    written = {}

    def imwrite(path, arr):
        ok, buf = cv2.imencode(".png", arr)                       # cv2's own float64 -> 8-bit conversion and PNG encoder
        assert ok
        written[os.path.basename(path).split("_", 1)[1][:-4]] = cv2.imdecode(buf, cv2.IMREAD_UNCHANGED)
        return True

    mask = torch.from_numpy(inputs["mask_u8"].astype(np.float64).reshape(H, W, 1)) / 255.0          # TESTB:569
    m3 = np.zeros((H, W, 3))
    for c in range(3):
        m3[:, :, c] = mask.numpy().reshape(H, W)                                                   # TESTB:578-582
    ns = dict(
        np=np, cv2=types.SimpleNamespace(imwrite=imwrite), model=types.SimpleNamespace(batch_size=1, img_height=H, img_width=W),
        rendered_images=torch.from_numpy(inputs["rendered"]), albedo=torch.from_numpy(inputs["albedo"]),
        depth=torch.from_numpy(inputs["depth"]), final_shading=torch.from_numpy(inputs["final"]),
        surface_normals=torch.from_numpy(inputs["normals"]), shadow_mask_weights=torch.from_numpy(inputs["shadow"]),
        training_images=inputs["image"].copy(), batch_list=np.arange(1), j=0, curr_img_names=["x.png"],
        curr_mask_fill_nose=mask, curr_mask_fill_nose_3_channels=m3)
    exec(compile(block, SCRIPT + ":%d-%d" % (FIRST, LAST), "exec"), ns)
    assert sorted(written) == ["albedo", "depth", "rendered_image", "shading", "shadow_mask", "surface_normals"], sorted(written)
    np.savez_compressed(os.path.join(OUT, "planes.npz"), lines=np.array([FIRST, LAST]),
                        **{"in_" + k: v for k, v in inputs.items()}, **{"out_" + k: v for k, v in written.items()})
    print({k: (v.shape, v.dtype) for k, v in written.items()}, os.path.getsize(os.path.join(OUT, "planes.npz")))


if __name__ == "__main__":
    main()
