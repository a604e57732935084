"""geomconsistentfr_b200 — B200 (sm_100a) implementation of the relight hot path of
andrewhou1/GeomConsistentFR: ray-march shadow mask, Lambertian shading/render, RelightNet CNN.

    csrc/      CUDA kernels + the C ABI (include/gfr_b200.h) -> csrc/libgfr_b200.so
    _lib.py    ctypes binding (no fallback: raises if the library is missing)
    ops.py     operator wrappers over the C ABI
    relightnet.py  drop-in RelightNet (reference constructor attrs, state_dict keys, forward signatures)
    inference.py   the TEST1 / TESTB driver bodies (composite, 8-bit export, border fix) around the model call
"""
from . import _lib, ops  # noqa: F401
from .relightnet import RelightNet, intrinsic_matrix  # noqa: F401
from .runner import RelightRunner  # noqa: F401
from .patchgan import PatchGAN  # noqa: F401
from .inference import lighting_transfer, relight, relight_single_image  # noqa: F401
from .lpips_metric import LPIPSAlex, masked_lpips  # noqa: F401
from .autograd import ShadowMarch, ShadeRender, SSIMPlanes, MaskedLosses, FlatAdam, dssim_loss  # noqa: F401

__version__ = "0.1.0"
