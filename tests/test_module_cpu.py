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
