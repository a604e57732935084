"""GPU: the lighting-transfer variant in TRAIN mode (= train_lighting_transfer.py, TRAIN_LT) against outputs and autograd
gradients of the UNMODIFIED script (tests/golden/lt_train.npz, made by oracle/make_golden_lt.py)."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def ffhq():
    return np.load(os.path.join(G, "ffhq.npz"))


def test_train_mode_matches_the_lighting_transfer_training_script(ffhq):
    """RelightNet(variant="lighting_transfer").train() = train_lighting_transfer.py's forward (TRAIN_LT: TRAIN's constants,
    1x1 bias-free shortcuts): outputs and autograd gradients for B = 3 vs the unmodified script (tests/golden/lt_train.npz).
    Tolerances = those of the default variant's train-mode test (tests/test_gpu_train.py)."""
    from geomconsistentfr_b200 import RelightNet, dssim_loss, intrinsic_matrix
    torch.backends.cudnn.allow_tf32 = False
    g = np.load(os.path.join(G, "lt_train.npz"))
    tnet = RelightNet(batch_size=3, variant="lighting_transfer")
    tnet.load_state_dict(torch.load(os.path.join(G, "model_epoch106.pth"), map_location="cpu"), strict=True)
    tnet = tnet.float().cuda().train()
    x = torch.from_numpy(ffhq["q"][g["sel"]] / 1020.0).float().cuda()
    mt = torch.from_numpy(g["masks01"].astype(np.float64)).view(3, 256, 256, 1).cuda()
    out = tnet(x, 200, intrinsic_matrix().cuda(), mt)
    assert len(out) == 8
    albedo, depth, shadow, _, _, rendered, unit_l, amb_v = out
    depth.retain_grad()
    assert np.abs(depth.detach().cpu().numpy() - g["depth"]).max() <= 2e-2
    assert np.abs(shadow.detach().cpu().numpy() - g["shadow"].astype(np.float32)).max() <= 5e-3
    assert np.abs(rendered.detach().cpu().numpy() - g["rendered"].astype(np.float32)).max() <= 5e-3
    assert np.abs(albedo.detach().cpu().numpy().mean(axis=(2, 3)) - g["albedo_mean"]).max() <= 5e-5
    assert np.abs(unit_l.detach().cpu().numpy().reshape(3, 3) - g["unit_light"]).max() <= 2e-5
    assert np.abs(amb_v.detach().cpu().numpy().reshape(3) - g["ambient"]).max() <= 2e-5
    m3 = mt.permute(0, 3, 1, 2).repeat(1, 3, 1, 1).float()
    target = x.permute(0, 3, 1, 2).contiguous()
    comp = rendered * m3 + (1.0 - m3) * target
    loss = 20.0 * torch.sum((rendered * m3 - target * m3) ** 2) / torch.sum(m3) + dssim_loss(comp, target)
    assert abs(float(loss) - float(g["loss"])) <= 2e-4 * float(g["loss"])
    loss.backward()
    gd, gd_ref = depth.grad.cpu().numpy(), g["grad_depth"].astype(np.float32)
    assert np.abs(gd - gd_ref).sum() / np.abs(gd_ref).sum() <= 2e-2
    for name, p in (("grad_shortcut_h1", tnet.conv_shortcut_h1_out.weight), ("grad_shortcut_h3", tnet.conv_shortcut_h3_out.weight),
                    ("grad_deconv_depth_shortcut_h6", tnet.deconv_depth_shortcut_h6_out.weight),
                    ("grad_sl2_w", tnet.linear_SL2.weight), ("grad_depth_head_w", tnet.conv_depth_c2_o.weight)):
        ref = g[name]
        assert p.grad.shape == ref.shape, name                    # [Cout, Cin, 1, 1]: the centre tap's gradient
        err = np.abs(p.grad.cpu().numpy() - ref).max() / np.abs(ref).max()
        assert err <= 2e-2, (name, err)
    for n_, p_ in tnet.named_parameters():
        assert p_.grad is not None and torch.isfinite(p_.grad).all(), n_
