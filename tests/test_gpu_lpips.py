"""GPU: LPIPS-Alex spatial (lpips/lpips.py:112-144, test_network.py:39-45) on the library's kernels against the reference's
VENDORED lpips module run unmodified (tests/golden/lpips.npz, oracle/make_golden_lpips.py): same seeded random AlexNet trunk
(the ImageNet weights are not available offline), the SHIPPED linear heads; spatial map, masked metric and the gradient of the
metric w.r.t. the predicted image ("LPIPS backward")."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    try:                                                   # the newer per-operator switch, where it exists
        torch.backends.cudnn.conv.fp32_precision = "ieee"
    except Exception:
        pass


def _model():
    from geomconsistentfr_b200.lpips_metric import LPIPSAlex
    from oracle.make_golden_lpips import trunk_init
    f = np.load(os.path.join(G, "lpips.npz"))
    m = LPIPSAlex()
    heads = {"lin%d.model.1.weight" % k: torch.from_numpy(f["lin%d" % k]).view(1, -1, 1, 1) for k in range(5)}
    missing, unexpected = m.load_state_dict(heads, strict=False)          # what lpips.py:109 does with alex.pth
    assert not unexpected and all(k.startswith(("net.", "scaling_layer.", "lins.")) for k in missing)
    trunk_init(m.net)
    return m.cuda(), f


def test_spatial_map_metric_and_gradient_vs_vendored_lpips():
    from geomconsistentfr_b200.lpips_metric import masked_lpips
    from oracle.make_golden_lpips import case
    m, f = _model()
    ref, pred, mask = (t.cuda() for t in case())
    pred = pred.clone().requires_grad_()
    ex = m(ref, pred)
    assert ex.shape == (2, 1, 256, 256)
    want = f["map"]
    # (1) the metric's own arithmetic against the vendored formulas (lpips.py:125-131, 16-18) evaluated by torch on the SAME trunk
    #     features: this isolates the library's kernels from the cuDNN-vs-CPU differences of the trunk convolutions
    import torch.nn.functional as F
    with torch.no_grad():
        o0, o1 = m.net(m.scaling_layer(ref)), m.net(m.scaling_layer(pred))
        val = 0
        for k in range(5):
            n0 = o0[k] / (torch.sqrt(torch.sum(o0[k] ** 2, dim=1, keepdim=True)) + 1e-10)
            n1 = o1[k] / (torch.sqrt(torch.sum(o1[k] ** 2, dim=1, keepdim=True)) + 1e-10)
            d = F.conv2d((n0 - n1) ** 2, m.lins[k].model[1].weight)
            val = val + F.interpolate(d, size=(256, 256), mode="bilinear", align_corners=False)
    e_kernel = float((ex.detach() - val).abs().max()) / float(val.abs().max())
    e_fixture = float(np.abs(ex.detach().cpu().numpy() - want).max() / np.abs(want).max())
    e_trunk = float(np.abs(val.cpu().numpy() - want).max() / np.abs(want).max())
    print("lpips map: kernels vs torch formulas %.2e, vs fixture %.2e, torch-on-cuda vs fixture %.2e" % (e_kernel, e_fixture, e_trunk))
    assert e_kernel <= 2e-6, e_kernel
    # (2) against the reference's own run (trunk convolutions by cuDNN here, by the CPU there)
    assert e_fixture <= max(2e-3, 2 * e_trunk), (e_fixture, e_trunk)
    metric = masked_lpips(ex, mask)
    assert np.abs(metric.cpu().numpy() - f["metric"]).max() <= max(2e-4, 2 * e_trunk) * np.abs(f["metric"]).max()
    # gradient of the metric (sum over the two images) w.r.t. the predicted image, through the kernels' backward
    cnt = torch.stack([(mask * ex[i, 0] > 0).sum() for i in range(2)]).float()
    (torch.stack([(mask * ex[i, 0]).sum() for i in range(2)]) / cnt).sum().backward()
    g, gw = pred.grad.cpu().numpy()[:, :, ::2, ::2], f["grad_pred_s2"]
    assert np.abs(g - gw).sum() / np.abs(gw).sum() <= 3e-2, np.abs(g - gw).sum() / np.abs(gw).sum()      # through cuDNN's trunk backward


def test_lpips_properties():
    m, _ = _model()
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.rand(1, 3, 128, 96, device="cuda", generator=g) * 2 - 1
    y = torch.rand(1, 3, 128, 96, device="cuda", generator=g) * 2 - 1
    assert float(m(x, x).abs().max()) <= 1e-12                                    # identical images: distance 0 (two trunk passes: cuDNN run-to-run rounding)
    assert torch.allclose(m(x, y), m(y, x), atol=1e-7)                            # symmetric
    assert float(m(x, y).min()) >= 0.0                                            # the shipped heads are non-negative
    assert torch.allclose(m((x + 1) / 2, (y + 1) / 2, normalize=True), m(x, y), atol=1e-6)


def test_layer_distance_and_upsample_backward_vs_torch_autograd():
    """The two differentiable LPIPS kernels alone (no trunk): forward and BOTH input gradients of the layer distance, and the
    bilinear upsample-accumulate, against torch autograd on the vendored formulas (lpips.py:125-131, 16-18) in float64."""
    import torch.nn.functional as F
    from geomconsistentfr_b200.lpips_metric import _LayerDistance, _UpsampleAdd
    g = torch.Generator(device="cuda").manual_seed(2)
    for (N, C, h, w, H, W) in ((2, 64, 15, 15, 64, 64), (1, 192, 7, 9, 40, 56)):
        f0 = torch.rand(N, C, h, w, device="cuda", generator=g).requires_grad_()
        f1 = torch.rand(N, C, h, w, device="cuda", generator=g).requires_grad_()
        lw = torch.rand(1, C, 1, 1, device="cuda", generator=g)
        G = torch.randn(N, H, W, device="cuda", generator=g)
        out = _UpsampleAdd.apply(torch.zeros(N, H, W, device="cuda"), _LayerDistance.apply(f0, f1, lw))
        (out * G).sum().backward()
        a, b = f0.detach().double().requires_grad_(), f1.detach().double().requires_grad_()
        n0 = a / (torch.sqrt(torch.sum(a ** 2, dim=1, keepdim=True)) + 1e-10)
        n1 = b / (torch.sqrt(torch.sum(b ** 2, dim=1, keepdim=True)) + 1e-10)
        ref = F.interpolate(F.conv2d((n0 - n1) ** 2, lw.double()), size=(H, W), mode="bilinear", align_corners=False)[:, 0]
        (ref * G.double()).sum().backward()
        assert float((out.double() - ref).abs().max()) <= 2e-6 * float(ref.abs().max())
        for mine, want in ((f0.grad, a.grad), (f1.grad, b.grad)):
            assert float((mine.double() - want).abs().max()) <= 2e-5 * float(want.abs().max())
