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
