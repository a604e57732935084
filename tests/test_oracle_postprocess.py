"""CPU: pins oracle/postprocess_oracle.py against the reference's own export code — tests/golden/planes.npz holds what the
UNMODIFIED source lines test_raytracing_relighting_CelebAHQ_DSSIM_8x.py:584-608 hand to cv2.imwrite (passed through cv2's
real float -> 8-bit PNG encoder) for synthetic forward outputs (oracle/make_golden_planes.py)."""
import os

import numpy as np

from oracle import postprocess_oracle as P

G = os.path.join(os.path.dirname(__file__), "golden")


def test_export_block_matches_the_reference_lines_bit_for_bit():
    f = np.load(os.path.join(G, "planes.npz"))
    assert list(f["lines"]) == [584, 608]
    mask = f["in_mask_u8"]
    comp = P.composite_bgr_u8(f["in_image"][0], f["in_rendered"][0], mask)
    assert np.array_equal(comp, f["out_rendered_image"])                      # TESTB:598-602, ties and saturation included
    planes = P.export_planes_u8(f["in_albedo"][0], f["in_depth"][0], f["in_shadow"][0], f["in_final"][0], f["in_normals"][0], mask)
    for k in ("shadow_mask", "albedo", "depth", "shading", "surface_normals"):   # TESTB:603-607
        assert planes[k].shape == f["out_" + k].shape, k
        assert np.array_equal(planes[k], f["out_" + k]), (k, int(np.abs(planes[k].astype(int) - f["out_" + k].astype(int)).max()))
