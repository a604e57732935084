"""TEST INFRASTRUCTURE — not product code.

Loads the UNMODIFIED reference scripts from /root/reference through dependency
shims so that the oracle restatement (oracle/relight_oracle.py) can be pinned
against the reference itself and so that golden fixtures can be generated
(oracle/make_golden.py).  /root/reference only exists in the authoring
container; nothing that runs on the GPU box imports this file.

The reference needs five shims to import on this image:
  1. Tensor.cuda / Module.cuda -> identity  (no GPU in the authoring container)
  2. np.asscalar                            (removed from NumPy; used at TRAIN:380-381, TEST1:357-358)
  3. kornia.geometry.depth.depth_to_normals (kornia 0.4.1 is not installed; call sites TRAIN:353, TEST1:326)
  4. pytorch_msssim.ssim                    (not installed, unpinned upstream; call site TRAIN:643)
  5. imageio.imread                         (not installed; cv2-backed, RGB order)

Shims 3 and 4 are restated from the published upstream algorithms
(kornia 0.4.1 `depth_to_normals`, pytorch_msssim `ssim`); their bodies live in
oracle/relight_oracle.py so the oracle and the shimmed reference share one
statement of the two un-vendored dependencies.
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch
import torch.nn as nn

_HERE = os.path.dirname(os.path.abspath(__file__))
_STAGED = os.path.join(_HERE, "_ref")                 # byte copies made by oracle/fetch_ref.py (git-ignored; ships to the GPU box)
_GOLDEN = os.path.join(_HERE, "..", "tests", "golden")


def _default_root():
    if os.environ.get("GFR_REFERENCE_ROOT"):
        return os.environ["GFR_REFERENCE_ROOT"]
    if os.path.isfile(os.path.join("/root/reference", "test_relight_single_image.py")):
        return "/root/reference"
    return _STAGED


REFERENCE_ROOT = _default_root()

_SCRIPTS = {
    "TRAIN": "train_raytracing_relighting_CelebAHQ_DSSIM_8x.py",
    "TEST1": "test_relight_single_image.py",
    "TESTB": "test_raytracing_relighting_CelebAHQ_DSSIM_8x.py",
    "TEST_LT": "test_relight_single_image_lighting_transfer.py",
    "TRAIN_LT": "train_lighting_transfer.py",
}


def reference_available():
    return os.path.isfile(os.path.join(REFERENCE_ROOT, _SCRIPTS["TEST1"]))


def _install_shims(cuda_identity=True):
    from oracle import relight_oracle as spec

    # 1. .cuda() -> identity (CPU runs only; with cuda_identity=False the reference's own .cuda() calls stay real —
    #    bench.py's `reference_gpu` leg on the GPU box)
    if cuda_identity:
        torch.Tensor.cuda = lambda self, *a, **k: self
        nn.Module.cuda = lambda self, *a, **k: self
    # 2. np.asscalar
    if not hasattr(np, "asscalar"):
        np.asscalar = lambda a: a.item()
    # 3. kornia
    if "kornia" not in sys.modules:
        kornia = types.ModuleType("kornia")
        geometry = types.ModuleType("kornia.geometry")
        depth = types.ModuleType("kornia.geometry.depth")
        depth.depth_to_normals = spec.depth_to_normals
        geometry.depth = depth
        kornia.geometry = geometry
        sys.modules["kornia"] = kornia
        sys.modules["kornia.geometry"] = geometry
        sys.modules["kornia.geometry.depth"] = depth
    # 4. pytorch_msssim
    if "pytorch_msssim" not in sys.modules:
        m = types.ModuleType("pytorch_msssim")
        m.ssim = spec.ssim
        m.ms_ssim = m.SSIM = m.MS_SSIM = None
        sys.modules["pytorch_msssim"] = m
    # 5. imageio
    if "imageio" not in sys.modules:
        import cv2

        def imread(path):
            img = cv2.imread(path, cv2.IMREAD_UNCHANGED)
            if img is None:
                raise FileNotFoundError(path)
            if img.ndim == 3:
                img = img[:, :, ::-1].copy()
            return img

        m = types.ModuleType("imageio")
        m.imread = imread
        sys.modules["imageio"] = m


_loaded = {}


def load_reference(short, cuda_identity=True):
    """Import one of the reference scripts (TRAIN / TEST1 / TESTB) unmodified."""
    if short in _loaded:
        return _loaded[short]
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    _install_shims(cuda_identity)
    path = os.path.join(REFERENCE_ROOT, _SCRIPTS[short])
    spec = importlib.util.spec_from_file_location("gfr_reference_" + short, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    _loaded[short] = mod
    return mod


def reference_model(short="TEST1", batch_size=None, weights=True, cuda_identity=True):
    """Instantiate the reference RelightNet (CPU).  `batch_size` patches the value the
    reference bakes into xx/yy at construction (TEST1:15,25-26 / TRAIN:41,52-53)."""
    mod = load_reference(short, cuda_identity)
    net = mod.RelightNet()
    if batch_size is not None and batch_size != net.batch_size:
        net.batch_size = batch_size
        net.xx = net.xx[:1].repeat(batch_size, 1, 1)
        net.yy = net.yy[:1].repeat(batch_size, 1, 1)
    if weights:
        rel = ("model_lighting_transfer", "model_epoch106.pth") if short in ("TEST_LT", "TRAIN_LT") else ("model", "model_epoch99.pth")
        path = os.path.join(REFERENCE_ROOT, *rel)
        if not os.path.isfile(path):               # the staged copy carries no weights: tests/golden holds byte-identical files
            path = os.path.join(_GOLDEN, rel[1])
        sd = torch.load(path, map_location="cpu")
        net.load_state_dict(sd)
    return net.float()
