"""GPU: the lighting-transfer variant (RelightNet(variant="lighting_transfer"), inference.lighting_transfer) against
outputs of the UNMODIFIED reference script test_relight_single_image_lighting_transfer.py (tests/golden/lt.npz, made by
oracle/make_golden_lt.py) — the two-pass transfer of TEST_LT:543-545 and both sides of the `+5` rule of TEST_LT:503."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def lt():
    return np.load(os.path.join(G, "lt.npz"))


@pytest.fixture(scope="module")
def ffhq():
    return np.load(os.path.join(G, "ffhq.npz"))


@pytest.fixture(scope="module")
def net():
    from geomconsistentfr_b200 import RelightNet
    n = RelightNet(variant="lighting_transfer")
    n.load_state_dict(torch.load(os.path.join(G, "model_epoch106.pth"), map_location="cpu"), strict=True)
    return n.cuda().eval()


def _inputs(ffhq, lt):
    names = list(ffhq["names"])
    i_in, i_ref = names.index(str(lt["input"])), names.index(str(lt["reference"]))
    return ffhq["q"][i_in] / 1020.0, ffhq["q"][i_ref] / 1020.0, ffhq["masks"][i_in]


@pytest.mark.parametrize("impl,prec", [("tc", 2), ("tc", 3), ("direct", 0)])
def test_two_pass_transfer_matches_the_reference(ffhq, lt, net, impl, prec):
    from geomconsistentfr_b200 import intrinsic_matrix
    net.cnn_impl, net.tc_precision = impl, (prec or net.tc_precision)
    try:
        x_in, x_ref, mask = _inputs(ffhq, lt)
        K = intrinsic_matrix(focal=700.0).cuda()
        m = (torch.from_numpy(mask).double() / 255.0).view(256, 256, 1).cuda()
        c = lambda a: torch.from_numpy(a).float()[None].cuda()
        with torch.no_grad():
            r1 = net(c(x_ref), 200, K, m, torch.zeros(1, 3, 1, 1).cuda(), torch.zeros(1, 1, 1).cuda())      # TEST_LT:543
            r2 = net(c(x_in), 200, K, m, r1[10], r1[11])                                                       # TEST_LT:545
        assert len(r1) == 12 and len(r2) == 12
        assert np.abs(r1[10].cpu().numpy().reshape(3) - lt["est_light"]).max() <= 2e-5
        assert np.abs(r1[11].cpu().numpy().reshape(1) - lt["est_ambient"]).max() <= 2e-5
        got = dict(depth=r2[1], shadow=r2[2], rendered=r2[5], unit_light=r2[6], ambient=r2[7], final=r2[8], normals=r2[9],
                   est_light_in=r2[10], est_ambient_in=r2[11])
        tol = dict(depth=5e-3, shadow=2e-3, rendered=2e-3, final=2e-3, normals=2e-4, unit_light=3e-5, ambient=3e-5,
                   est_light_in=3e-5, est_ambient_in=3e-5)            # the bars of the default variant (tests/test_gpu_relight.py)
        for k, v in got.items():
            assert np.abs(v.cpu().numpy() - lt["transfer_" + k]).max() <= tol[k], k
    finally:
        net.cnn_impl, net.tc_precision = "tc", 2


@pytest.mark.parametrize("tag", ["near", "far"])
def test_bonus_rule(ffhq, lt, net, tag):
    from geomconsistentfr_b200 import intrinsic_matrix
    x_in, _, mask = _inputs(ffhq, lt)
    m = (torch.from_numpy(mask).double() / 255.0).view(256, 256, 1).cuda()
    tl = torch.from_numpy(lt[tag + "_light"]).view(1, 3, 1, 1).cuda()
    with torch.no_grad():
        r = net(torch.from_numpy(x_in).float()[None].cuda(), 200, intrinsic_matrix(focal=700.0).cuda(), m, tl,
                torch.full((1, 1, 1), float(lt[tag + "_ambient"])).cuda())
    assert np.abs(r[2].cpu().numpy() - lt[tag + "_shadow"]).max() <= 2e-3
    inside = mask > 0
    s = r[2].cpu().numpy()[0][inside]
    assert (s.min() > 0.97) if tag == "near" else (s.min() < 0.5)


def test_driver_matches_oracle_postprocess_and_estimates(ffhq, lt, net):
    from geomconsistentfr_b200 import lighting_transfer
    from oracle import postprocess_oracle as P
    x_in, x_ref, mask = _inputs(ffhq, lt)
    res, est_l, est_a = lighting_transfer(net, x_in, x_ref, mask)
    assert np.abs(est_l - lt["est_light"]).max() <= 2e-5 and abs(est_a - float(lt["est_ambient"][0])) <= 2e-5
    assert set(res) == {"rendered_image", "shadow_mask", "albedo", "depth", "shading", "surface_normals"}
    want = P.composite_bgr_u8(x_in, lt["transfer_rendered"][0], mask)          # the reference's own rendered output, composited
    d = np.abs(res["rendered_image"][0].astype(np.int32) - want.astype(np.int32))
    assert d.max() <= 1 and d.mean() < 0.02
    sh = P.export_planes_u8(np.zeros((3, 256, 256), np.float32), lt["transfer_depth"][0], lt["transfer_shadow"][0],
                            lt["transfer_final"][0], lt["transfer_normals"][0], mask)
    for k in ("shadow_mask", "shading", "surface_normals", "depth"):
        assert np.abs(res[k][0].astype(np.int32) - sh[k].astype(np.int32)).max() <= 1, k


def test_wrong_variant_is_refused():
    from geomconsistentfr_b200 import RelightNet, lighting_transfer
    with pytest.raises(RuntimeError):
        lighting_transfer(RelightNet().cuda().eval(), np.zeros((256, 256, 3)), np.zeros((256, 256, 3)), np.zeros((256, 256), np.uint8))
