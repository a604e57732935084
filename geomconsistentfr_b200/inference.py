"""Inference drivers: the body of `main()` in test_relight_single_image.py (TEST1:507-620) and of the Multi-PIE batch
script (TESTB:542-608) around the model call, minus file I/O — arrays in, the 8-bit images `cv2.imwrite` would store out.

Differences from the reference, none of them visible in the results:
  * the forward's fp32 planes are composited / quantised on the device (csrc/postprocess.cu) and only u8 images are
    copied to the host (the reference copies every fp32 plane and does this in numpy);
  * any batch size (the reference bakes batch_size = 1 into `xx, yy`, TEST1:15,25-26);
  * optional `fix_border=True` applies fix_border_artifacts_CVPR2022.m (a separate MATLAB pass in the reference,
    README.md:54) on the device; with it the 10 shipped FFHQ_relighting_results/ PNGs are reproduced on every pixel.
PNG reading/writing stays with the caller (cv2 / imageio); `write_pngs` is a convenience for the six TESTB files."""
import os

import numpy as np
import torch

from . import ops
from .relightnet import intrinsic_matrix

SHIPPED_BORDER_MAX_SUM = 30      # `0 < boxsum <= 30` reproduces the shipped PNGs; the .m file's literal `< 30` is 29


def _as_batch(images, masks, lights):
    img = torch.as_tensor(np.asarray(images)) if not torch.is_tensor(images) else images
    if img.dim() == 3:
        img = img[None]
    B, H, W, _ = img.shape
    m = torch.as_tensor(np.asarray(masks)) if not torch.is_tensor(masks) else masks
    if m.dtype != torch.uint8:
        raise RuntimeError("masks must be uint8 skin masks as read from disk (values 0..255), not /255 floats")
    m = m.reshape(-1, H, W)
    if m.shape[0] not in (1, B):
        raise RuntimeError("masks must be [H,W] or [B,H,W]")
    L = torch.as_tensor(np.asarray(lights, dtype=np.float32)) if not torch.is_tensor(lights) else lights.float()
    L = L.reshape(-1, 3)
    if L.shape[0] == 1 and B > 1:
        L = L.expand(B, 3)
    if L.shape[0] != B:
        raise RuntimeError("lights must be [3] or [B,3]")
    return img, m, L.contiguous()


@torch.no_grad()
def relight(model, images, masks, lights, ambient=0.5, epoch=200, fix_border=False, border_max_sum=SHIPPED_BORDER_MAX_SUM,
            planes=False):
    """TEST1:582-620 for a batch.  images [B,256,256,3] (or one [256,256,3]) RGB in [0,1] — float64 as the reference
    holds them (`training_images`, TEST1:513-515) or float32; masks uint8 [256,256] (shared, one forward per distinct
    mask like the reference) or [B,256,256]; lights [B,3] | [3] target light directions (`training_lightings[:, 1:4]`);
    `ambient` is accepted and unused exactly like the reference's `target_ambient_values` (TEST1:169).

    Returns a dict of host uint8 numpy arrays: 'rendered_image' [B,H,W,3] BGR (what cv2.imwrite stores at TEST1:620 /
    TESTB:602) and, with planes=True, 'shadow_mask', 'albedo', 'depth', 'shading', 'surface_normals' (TESTB:603-607; depth is
    normalised per face like the reference's batch of one)."""
    if model.training:
        raise RuntimeError("relight() is the inference path: call model.eval() first (TEST1:512)")
    dev = model.device
    if dev.type != "cuda":
        raise RuntimeError("relight() needs the model on a CUDA device (there is no CPU path)")
    img, m, L = _as_batch(images, masks, lights)
    B, H, W, _ = img.shape
    img_d = img.to(dev, non_blocking=True)
    m_d = m.to(dev, non_blocking=True)
    K = intrinsic_matrix(H, W)
    x = img_d.float()
    tl = L.to(dev).view(B, 3, 1, 1)
    amb = torch.full((B, 1, 1), float(ambient), device=dev)
    shared = m_d.shape[0] == 1
    # the TEST1 forward takes ONE mask for its batch (TEST1:488); this module also accepts one per face
    # (the march only tests mask != 0, TEST1:488, so the u8 mask goes in as it is instead of the reference's mask / 255.0)
    o = model(x, epoch, K, m_d.view(-1, H, W, 1)[0 if shared else slice(None)], tl, amb, None)
    albedo, depth, shadow, rendered, final_shading, normals = o[0], o[1], o[2], o[5], o[8], o[9]
    mask_arg = m_d[0] if shared else m_d
    res = {"rendered_image": ops.composite_bgr_u8(img_d, rendered, mask_arg)}
    if planes:
        if B == 1:
            res.update(ops.export_planes_u8(albedo, depth, shadow, final_shading, normals, mask_arg))
        else:                   # TESTB normalises depth over its batch of ONE face: per-face range here
            per = [ops.export_planes_u8(albedo[b:b + 1], depth[b:b + 1], shadow[b:b + 1], final_shading[b:b + 1],
                                        normals[b:b + 1], m_d[0] if shared else m_d[b]) for b in range(B)]
            res.update({k: torch.cat([p[k] for p in per], 0) for k in per[0]})
    if fix_border:
        res["rendered_image"] = ops.border_median_fix_u8(res["rendered_image"], mask_arg, border_max_sum)
    host = {k: v.cpu().numpy() for k, v in res.items()}
    return host


@torch.no_grad()
def lighting_transfer(model, input_image, reference_image, mask, epoch=200, planes=True):
    """`main()` of test_relight_single_image_lighting_transfer.py (TEST_LT:516-579): pass 1 runs the REFERENCE image with a
    zero target light to estimate its light direction (z floored at 0.16) and ambient; pass 2 relights the INPUT image
    with them (f = 700, TEST_LT:530-531).  `model` is RelightNet(variant="lighting_transfer") in eval mode; images
    [256,256,3] RGB in [0,1] (f64 like the reference, or f32); mask uint8 [256,256] (the input's).
    Returns (dict of host uint8 images as `relight`, estimated unit light [3], estimated ambient)."""
    if getattr(model, "variant", "default") != "lighting_transfer":
        raise RuntimeError("lighting_transfer() needs RelightNet(variant='lighting_transfer')")
    if model.training:
        raise RuntimeError("lighting_transfer() is an inference path: call model.eval() first (TEST_LT:521)")
    dev = model.device
    img, m, _ = _as_batch(input_image, mask, (0.0, 0.0, 1.0))
    ref, _, _ = _as_batch(reference_image, mask, (0.0, 0.0, 1.0))
    B, H, W, _ = img.shape
    K = intrinsic_matrix(H, W, focal=model.focal_length)
    img_d, ref_d, m_d = img.to(dev), ref.to(dev), m.to(dev)
    mk = m_d.view(-1, H, W, 1)[0]
    zero = torch.zeros(B, 4, device=dev)
    r1 = model(ref_d.float(), epoch, K, mk, zero[:, 1:4].view(B, 3, 1, 1), zero[:, 0].view(B, 1, 1))            # TEST_LT:543
    est_l, est_a = r1[10], r1[11]
    o = model(img_d.float(), epoch, K, mk, est_l.view(B, 3, 1, 1), est_a.view(B, 1, 1))                          # TEST_LT:545
    res = {"rendered_image": ops.composite_bgr_u8(img_d, o[5], m_d[0])}
    if planes:
        res.update(ops.export_planes_u8(o[0], o[1], o[2], o[8], o[9], m_d[0]))                                   # TEST_LT:574-579
    return {k: v.cpu().numpy() for k, v in res.items()}, est_l.reshape(B, 3)[0].cpu().numpy(), float(est_a.reshape(B)[0])


def relight_single_image(model, image, mask, light, **kw):
    """TEST1 `main()` for one image: image [256,256,3] RGB in [0,1], mask uint8 [256,256], light (x,y,z).
    Returns the [256,256,3] uint8 BGR array the reference writes to FFHQ_relighting_results/<name>_rendered_image.png."""
    return relight(model, image, mask, light, **kw)["rendered_image"][0]


def write_pngs(out_dir, names, result):
    """TESTB:602-607 file names: <name>_rendered_image.png, _shadow_mask, _albedo, _depth, _shading, _surface_normals."""
    import cv2
    os.makedirs(out_dir, exist_ok=True)
    for i, n in enumerate(names):
        stem = os.path.splitext(n)[0]
        for k, v in result.items():
            cv2.imwrite(os.path.join(out_dir, "%s_%s.png" % (stem, k)), v[i])


def _main(argv=None):
    """Command-line form of the two single-image scripts (file I/O through cv2, which the reference also uses to write):

        python -m geomconsistentfr_b200.inference relight  MODEL.pth IMAGE MASK LX,LY,LZ OUT.png [--fix-border]
        python -m geomconsistentfr_b200.inference transfer MODEL.pth INPUT REFERENCE MASK OUT_DIR

    `relight` = test_relight_single_image.py main() (TEST1:507-620; the image is resized to 256x256 like TEST1:515);
    `transfer` = test_relight_single_image_lighting_transfer.py main() (TEST_LT:516-579; 256x256 inputs, six PNGs)."""
    import argparse
    import cv2
    from .relightnet import RelightNet
    ap = argparse.ArgumentParser(prog="python -m geomconsistentfr_b200.inference", description=_main.__doc__,
                                 formatter_class=argparse.RawDescriptionHelpFormatter)
    sub = ap.add_subparsers(dest="cmd", required=True)
    r = sub.add_parser("relight")
    r.add_argument("model"); r.add_argument("image"); r.add_argument("mask"); r.add_argument("light"); r.add_argument("out")
    r.add_argument("--fix-border", action="store_true")
    t = sub.add_parser("transfer")
    t.add_argument("model"); t.add_argument("input"); t.add_argument("reference"); t.add_argument("mask"); t.add_argument("out_dir")
    a = ap.parse_args(argv)

    def read_rgb01(path):
        img = cv2.imread(path, cv2.IMREAD_COLOR)
        if img is None:
            raise FileNotFoundError(path)
        img = img[:, :, ::-1] / 255.0                               # imageio.imread(...)/255.0, RGB (TEST1:515)
        return img if img.shape[:2] == (256, 256) else cv2.resize(img, (256, 256))

    def read_mask(path):
        m = cv2.imread(path, cv2.IMREAD_UNCHANGED)
        if m is None:
            raise FileNotFoundError(path)
        return np.ascontiguousarray(m[:, :, 0] if m.ndim == 3 else m).astype(np.uint8)

    net = RelightNet(variant="lighting_transfer" if a.cmd == "transfer" else "default")
    net.load_state_dict(torch.load(a.model, map_location="cpu"), strict=True)
    net = net.float().cuda().eval()
    if a.cmd == "relight":
        light = tuple(float(v) for v in a.light.split(","))
        out = relight_single_image(net, read_rgb01(a.image), read_mask(a.mask), light, fix_border=a.fix_border)
        cv2.imwrite(a.out, out)                                     # TEST1:620
    else:
        res, est_l, est_a = lighting_transfer(net, read_rgb01(a.input), read_rgb01(a.reference), read_mask(a.mask))
        write_pngs(a.out_dir, [os.path.basename(a.input)], res)    # TEST_LT:574-579
        print("estimated light", est_l.tolist(), "ambient", est_a)


if __name__ == "__main__":
    _main()
