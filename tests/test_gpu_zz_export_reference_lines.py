"""GPU: the device output stage against what the UNMODIFIED reference lines TESTB:584-608 hand to cv2.imwrite
(tests/golden/planes.npz, made by oracle/make_golden_planes.py)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


def test_export_kernels_match_the_reference_export_lines():
    """The device output stage against what the UNMODIFIED lines TESTB:584-608 hand to cv2.imwrite (tests/golden/planes.npz,
    made by oracle/make_golden_planes.py): bit-exact."""
    from geomconsistentfr_b200 import ops
    f = np.load(os.path.join(G, "planes.npz"))
    c = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    mask = c(f["in_mask_u8"])
    comp = ops.composite_bgr_u8(c(f["in_image"]), c(f["in_rendered"]), mask)
    assert np.array_equal(comp[0].cpu().numpy(), f["out_rendered_image"])
    planes = ops.export_planes_u8(c(f["in_albedo"]), c(f["in_depth"]), c(f["in_shadow"]), c(f["in_final"]), c(f["in_normals"]), mask)
    for k in ("shadow_mask", "albedo", "depth", "shading", "surface_normals"):
        assert np.array_equal(planes[k][0].cpu().numpy(), f["out_" + k]), k
