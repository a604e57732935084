"""CPU: the drop-in module has the reference's interface (names, attributes, state_dict keys)."""
import os

import pytest
import torch

from geomconsistentfr_b200 import RelightNet, intrinsic_matrix
from oracle import relight_oracle as O

G = os.path.join(os.path.dirname(__file__), "golden")


def test_state_dict_is_strictly_compatible_with_reference_weights():
    sd = torch.load(os.path.join(G, "model_epoch99.pth"), map_location="cpu")
    net = RelightNet()
    missing, unexpected = net.load_state_dict(sd, strict=True)
    assert not missing and not unexpected
    assert len(net.state_dict()) == 400 == len(sd)
    assert list(net.state_dict().keys()).sort() == list(sd.keys()).sort()
    assert set(net.state_dict()) == set(O.RelightNetOracle().state_dict())


def test_reference_constructor_attributes():
    net = RelightNet(batch_size=3)
    assert (net.batch_size, net.img_height, net.img_width, net.lr, net.df_dim) == (3, 256, 256, 0.0001, 64)
    assert (net.directional_intensity, net.light_distance, net.num_sample_points, net.GD_ratio) == (0.5, 4013.0, 160, 5)
    xx, yy = O.pixel_grid()
    assert torch.equal(net.xx[0], xx) and torch.equal(net.yy[2], yy) and net.xx.shape == (3, 256, 256)
    assert torch.equal(intrinsic_matrix(), O.intrinsic_matrix())


def test_no_cpu_fallback():
    net = RelightNet().eval()
    with pytest.raises(RuntimeError):
        net(torch.zeros(1, 256, 256, 3), 200, intrinsic_matrix(), torch.ones(256, 256, 1),
            torch.ones(1, 3, 1, 1), torch.ones(1, 1, 1), None)


def test_bench_workload_generator_matches_the_oracle_generator():
    """bench.py's product arm builds its inputs from the package (no oracle import); both arms must see the same faces."""
    from geomconsistentfr_b200 import synthetic as S
    assert S.LIGHTS_18 == O.LIGHTS_18
    for seed, H, W in ((0, 256, 256), (5, 64, 64), (17, 48, 80)):
        d0, m0 = S.synthetic_face(seed, H, W, noise=2.0)
        d1, m1 = O.synthetic_face(seed, H, W, noise=2.0)
        assert torch.equal(d0, d1) and torch.equal(m0, m1)
    L = torch.tensor(S.LIGHTS_18)
    assert torch.equal(S.light_point(L)[1], O.light_point(L)[1])


def test_product_code_never_imports_the_oracle():
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "geomconsistentfr_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle", src, re.M), f


def test_lighting_transfer_variant_loads_the_reference_weights_strictly():
    """TEST_LT:36-42,66-76,119-129: nine 1x1 bias-free shortcuts; model_lighting_transfer/model_epoch106.pth has 391 tensors."""
    sd = torch.load(os.path.join(G, "model_epoch106.pth"), map_location="cpu")
    net = RelightNet(variant="lighting_transfer")
    missing, unexpected = net.load_state_dict(sd, strict=True)
    assert not missing and not unexpected and len(sd) == 391 == len(net.state_dict())
    assert net.conv_shortcut_h3_out.weight.shape == (155, 64, 1, 1) and net.conv_shortcut_h3_out.bias is None
    assert net.deconv_depth_shortcut_all_features.weight.shape == (128, 64, 1, 1)
    assert (net.directional_intensity, net.num_sample_points, net.depth_offset, net.focal_length) == (0.41, 159, 1410.0, 700.0)
    import numpy as np
    assert np.array_equal(net.sample_table(), np.arange(0.03, 0.825, 0.005))              # TEST_LT:451
    assert np.array_equal(RelightNet().sample_table(), np.arange(0.025, 0.825, 0.005))    # TRAIN:468
    assert set(net.state_dict()) == set(O.RelightNetOracle(variant="lighting_transfer").state_dict())
    with pytest.raises(ValueError):
        RelightNet(variant="nope")
    # folded eval weights: the 1x1 shortcut becomes the centre tap of a 3x3 kernel with a zero bias
    f = net.eval()._folded_weights()
    w, b = f["conv_shortcut_h1_out"]
    assert w.shape == (32, 16, 3, 3) and float(w[:, :, 0, 0].abs().max()) == 0.0 and float(w[:, :, 1, 1].abs().max()) > 0.0
    w, b = f["deconv_albedo_shortcut_h5_out"]
    assert w.shape == (32, 64, 3, 3) and float(w[:, :, 2, 1].abs().max()) == 0.0


def test_inference_driver_argument_shapes():
    """inference._as_batch: one image or a batch, shared or per-face uint8 masks, one light or one per face."""
    import numpy as np
    from geomconsistentfr_b200.inference import _as_batch
    img, m, L = _as_batch(np.zeros((256, 256, 3)), np.zeros((256, 256), np.uint8), (0.0, 0.6, 0.8))
    assert img.shape == (1, 256, 256, 3) and img.dtype == torch.float64 and m.shape == (1, 256, 256) and L.shape == (1, 3)
    img, m, L = _as_batch(np.zeros((4, 64, 96, 3), np.float32), np.zeros((4, 64, 96), np.uint8), (0.0, 0.6, 0.8))
    assert img.shape == (4, 64, 96, 3) and m.shape == (4, 64, 96) and L.shape == (4, 3) and L.is_contiguous()
    with pytest.raises(RuntimeError):
        _as_batch(np.zeros((2, 64, 96, 3)), np.zeros((3, 64, 96), np.uint8), (0.0, 0.6, 0.8))        # 3 masks for 2 faces
    with pytest.raises(RuntimeError):
        _as_batch(np.zeros((2, 64, 96, 3)), np.zeros((64, 96), np.float32), (0.0, 0.6, 0.8))           # /255 float mask
    with pytest.raises(RuntimeError):
        _as_batch(np.zeros((2, 64, 96, 3)), np.zeros((64, 96), np.uint8), np.zeros((3, 3)))            # 3 lights for 2 faces
