"""GPU parity of the full relight forward (CNN + normals + march + render) through the drop-in module:
vs the reference's own outputs (tests/golden), vs the 10 PNGs the reference ships, vs the CPU oracle."""
import os

import numpy as np
import pytest
import torch

from oracle import relight_oracle as O

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def ffhq():
    return np.load(os.path.join(G, "ffhq.npz"))


@pytest.fixture(scope="module")
def net():
    from geomconsistentfr_b200 import RelightNet
    n = RelightNet()
    n.load_state_dict(torch.load(os.path.join(G, "model_epoch99.pth"), map_location="cpu"), strict=True)
    return n.float().cuda().eval()


@pytest.fixture(scope="module")
def oracle_net():
    n = O.RelightNetOracle()
    n.load_state_dict(torch.load(os.path.join(G, "model_epoch99.pth"), map_location="cpu"), strict=True)
    return n.eval()


def _inputs(ffhq, idx):
    x = torch.from_numpy(ffhq["q"][idx] / 1020.0).float()
    if x.dim() == 3:
        x = x[None]
    return x


@pytest.mark.parametrize("impl", ["tc", "tc-f16", "direct"])
@pytest.mark.parametrize("epoch", [200, 11, 0])
def test_cnn_vs_oracle(net, oracle_net, ffhq, epoch, impl):
    """CNN (BN folded; "tc" = tcgen05 3xTF32 convs, "direct" = exact-fp32 CUDA-core convs) vs torch fp32 on CPU, all
    epoch gates (TRAIN:245,258,271,283).
    Tolerances: albedo (sigmoid, [0,1]) 2e-5; light head 2e-5; depth = 100 x head, |depth| ~ 150: 5e-3."""
    x = _inputs(ffhq, [0, 5])
    with torch.no_grad():
        a_ref, d_ref, sl_ref = oracle_net.cnn(x, epoch)
        net.cnn_impl, net.tc_precision = ("tc", 2) if impl == "tc-f16" else (impl, 3)      # "tc" = 3xTF32
        try:
            a, d, sl = net._cnn_eval(x.cuda(), epoch)
        finally:
            net.cnn_impl, net.tc_precision = "tc", 2
    assert (a.cpu() - a_ref).abs().max() <= 2e-5
    assert (sl.cpu() - sl_ref).abs().max() <= 2e-5
    assert (d.cpu() - d_ref).abs().max() <= 5e-3


def test_forward_vs_reference_outputs(net, ffhq):
    """Every output of the reference forward (TEST1:505) for 00508, computed by the UNMODIFIED reference.
    Shadow/rendered tolerance 2e-3: the depth map feeding the march differs by ~1e-3 (fp32 conv order)."""
    from geomconsistentfr_b200 import intrinsic_matrix
    names = list(ffhq["names"])
    i = names.index("00508")
    x = _inputs(ffhq, i)
    m = torch.from_numpy(ffhq["masks"][i].astype(np.float64).reshape(256, 256, 1)) / 255.0
    tl = torch.from_numpy(ffhq["lights"][i]).view(1, 3, 1, 1)
    o = net(x.cuda(), 200, intrinsic_matrix().cuda(), m.cuda(), tl.cuda(), torch.full((1, 1, 1), 0.5).cuda(), None)
    assert len(o) == 10
    got = dict(albedo=o[0], depth=o[1], shadow=o[2], rendered=o[5], ambient=o[7], final=o[8], normals=o[9])
    tol = dict(albedo=2e-5, depth=5e-3, shadow=2e-3, rendered=2e-3, ambient=2e-5, final=2e-3, normals=2e-4)
    for k, v in got.items():
        diff = np.abs(v.cpu().numpy() - ffhq["00508_" + k])
        assert diff.max() <= tol[k], (k, diff.max())
    assert o[3].shape == (1, 256, 256) and o[4].shape == (1, 256, 256) and o[6].shape == (1, 3, 1, 1)


def _interior(mask_u8):
    m = (mask_u8 > 0).astype(np.float32)
    box = torch.nn.functional.avg_pool2d(torch.from_numpy(m)[None, None], 7, 1, 3, divisor_override=1)[0, 0].numpy()
    return box == 49


@pytest.mark.parametrize("precision", [3, 2])
def test_ten_shipped_pngs_batched(net, ffhq, precision):
    """The reference's known-answer vectors, FFHQ_relighting_results/*.png (TEST1:614-620 composite), mask
    interior, <= 1 grey level.  All 10 faces go through ONE batched forward per distinct light (the
    reference runs B=1)."""
    from geomconsistentfr_b200 import intrinsic_matrix
    K = intrinsic_matrix().cuda()
    net.tc_precision = precision
    for i in range(10):
        x = _inputs(ffhq, i)
        m = torch.from_numpy(ffhq["masks"][i].astype(np.float64).reshape(256, 256, 1)) / 255.0
        tl = torch.from_numpy(ffhq["lights"][i]).view(1, 3, 1, 1)
        o = net(x.cuda(), 200, K, m.cuda(), tl.cuda(), torch.full((1, 1, 1), 0.5).cuda(), None)
        rendered = o[5][0].permute(1, 2, 0).cpu().numpy().astype(np.float64)
        mask01 = ffhq["masks"][i].astype(np.float64)[..., None] / 255.0
        bgr = np.clip(np.rint(255.0 * rendered[:, :, ::-1] * mask01), 0, 255)
        inside = _interior(ffhq["masks"][i])
        diff = np.abs(bgr - ffhq["pngs_bgr"][i].astype(np.float64))[inside]
        assert diff.max() <= 1.0, (str(ffhq["names"][i]), diff.max())
        assert diff.mean() < 0.02
    net.tc_precision = 2


def test_train_signature_in_eval_mode_vs_oracle(net, oracle_net, ffhq):
    """TRAIN:196 signature (predicted light with z clamp, per-image masks, no +5 bonus, 8-tuple) in eval mode."""
    from geomconsistentfr_b200 import intrinsic_matrix
    idx = [1, 4, 3]
    x = _inputs(ffhq, idx)
    masks = torch.from_numpy((ffhq["masks"][idx] > 128).astype(np.float64)).view(3, 256, 256, 1)
    with torch.no_grad():
        ref = oracle_net.forward_train(x, 200, O.intrinsic_matrix(), masks)
    o = net(x.cuda(), 200, intrinsic_matrix().cuda(), masks.cuda())
    assert len(o) == 8
    for k, (a, b) in enumerate(zip(o, ref)):
        tol = 5e-3 if k == 1 else 2e-3
        assert a.shape == b.shape, k
        assert (a.cpu() - b).abs().max() <= tol, (k, float((a.cpu() - b).abs().max()))


def test_relight_sweep_equals_per_light_forward(net, ffhq):
    """One CNN pass + L lights (relight_sweep) == L full forwards (TESTB:565-583 semantics), bit for bit."""
    from geomconsistentfr_b200 import intrinsic_matrix
    K = intrinsic_matrix().cuda()
    x = _inputs(ffhq, [2, 7]).cuda()
    m = torch.from_numpy(ffhq["masks"][2].astype(np.float64).reshape(256, 256, 1)).cuda() / 255.0
    lights = torch.tensor(O.LIGHTS_18[3:6], dtype=torch.float32)
    sw = net.relight_sweep(x, 200, K, m, lights)
    assert sw["rendered"].shape == (2, 3, 3, 256, 256)
    for j in range(3):
        tl = lights[j].view(1, 3, 1, 1).expand(2, 3, 1, 1).contiguous().cuda()
        o = net(x, 200, K, m, tl, torch.full((2, 1, 1), 0.5).cuda(), None)
        assert torch.equal(sw["rendered"][:, j], o[5])
        assert torch.equal(sw["shadow"][:, j], o[2])


def test_p16_fused_decoder_tail_equals_the_two_launch_tail(net, ffhq):
    """The P16 path's default (gfr_conv3x3_p16_head_fwd: c2_1 with the 1x1 tail in its epilogue) against its two-launch form on the
    shipped weights: the only difference is that the 16-channel activation between them is not rounded to the fp16 pair
    (22 bits), so albedo agrees to 1e-6 and depth (100 x a value of order 1) to 2e-4."""
    x = _inputs(ffhq, [1, 4, 7]).cuda()
    assert net.tc_precision == 2 and net.p16 and net.fuse_head_p16
    try:
        with torch.no_grad():
            a1, d1, sl1 = net._cnn_eval(x, 200)
            net.fuse_head_p16 = False
            a0, d0, sl0 = net._cnn_eval(x, 200)
    finally:
        net.fuse_head_p16 = True
    assert float((a1 - a0).abs().max()) <= 1e-6, float((a1 - a0).abs().max())
    assert float((d1 - d0).abs().max()) <= 2e-4, float((d1 - d0).abs().max())
    assert torch.equal(sl1, sl0)


@pytest.mark.parametrize("precision", [2, 3])
def test_fused_decoder_tail_equals_the_two_launch_tail(net, ffhq, precision):
    """gfr_conv3x3_tc_head_fwd (c2_1 with c2_2 -> c2_3 -> c2_o in its epilogue) against gfr_conv3x3_tc_fwd followed by
    gfr_head_1x1_fwd: same accumulation order, so albedo and depth agree to the last bit or two."""
    x = _inputs(ffhq, [1, 4, 7]).cuda()
    net.tc_precision, net.p16 = precision, False         # the first-generation kernel (conv_tc.cu) in both precisions
    try:
        with torch.no_grad():
            net.fuse_head = True
            a1, d1, sl1 = net._cnn_eval(x, 200)
            net.fuse_head = False
            a0, d0, sl0 = net._cnn_eval(x, 200)
    finally:
        net.fuse_head, net.tc_precision, net.p16 = False, 2, True
    assert a1.shape == a0.shape == (3, 3, 256, 256) and d1.shape == d0.shape == (3, 1, 256, 256)
    assert float((a1 - a0).abs().max()) <= 2e-7
    assert float((d1 - d0).abs().max()) <= 2e-5          # 100 x a value of order 1
    assert torch.equal(sl1, sl0)
