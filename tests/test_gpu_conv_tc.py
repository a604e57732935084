"""GPU parity of the tcgen05 (3xTF32) 3x3 convolution and the C4-layout helpers, through the C ABI, against a plain
torch fp32 reference (cuDNN with TF32 disabled)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def test_c4_roundtrip():
    from geomconsistentfr_b200 import ops
    x = torch.randn(2, 155, 16, 16, device="cuda")
    c4 = ops.nchw_to_c4(x)
    assert c4.data.shape == (2, 39, 16, 16, 4)
    assert torch.equal(c4.data[:, 38, :, :, 3], torch.zeros(2, 16, 16, device="cuda"))      # channel padding is zero
    assert torch.equal(ops.c4_to_nchw(c4), x)
    p = ops.maxpool2_c4_fwd(c4)
    assert torch.equal(ops.c4_to_nchw(p), F.max_pool2d(x, 2))
    u = ops.upsample2_c4_fwd(p, add=c4)
    assert torch.equal(ops.c4_to_nchw(u), F.interpolate(F.max_pool2d(x, 2), scale_factor=2, mode="nearest") + x)


@pytest.mark.parametrize("N,Cin,Cout,H,W,NT", [
    (2, 16, 16, 32, 32, 16),
    (1, 16, 16, 256, 256, 16),
    (2, 32, 32, 64, 64, 32),
    (2, 16, 32, 64, 64, 16),
    (1, 64, 64, 32, 32, 64),
    (2, 64, 155, 16, 16, 32),
    (1, 155, 155, 16, 16, 32),
    (2, 128, 64, 16, 16, 16),
])
def test_conv3x3_tc_vs_torch(N, Cin, Cout, H, W, NT):
    """Tolerance: 3xTF32 keeps ~2^-21 relative error per product; |out| ~ sqrt(9*Cin) -> 2e-5 * scale absolute."""
    from geomconsistentfr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(N * 1000 + Cin + Cout + H)
    x = torch.randn(N, Cin, H, W, device="cuda", generator=g)
    w = torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / (3.0 * Cin ** 0.5)
    b = torch.randn(Cout, device="cuda", generator=g)
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=1)
    out = ops.conv3x3_tc_fwd(ops.nchw_to_c4(x), ops.conv_tc_pack_weights(w, NT), b, Cout, NT, act=None)
    got = ops.c4_to_nchw(out)
    err = (got.double() - ref).abs().max().item()
    ref32 = (F.conv2d(x, w, b, padding=1).double() - ref).abs().max().item()
    assert err <= max(4.0 * ref32, 2e-6), (err, ref32)


def test_conv3x3_tc_epilogue():
    """out = scale * (lrelu(conv + bias + res) + up2(post)) — the fused skip-block epilogue (TRAIN:240-246)."""
    from geomconsistentfr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(7)
    N, C, H, W = 2, 32, 64, 64
    x = torch.randn(N, C, H, W, device="cuda", generator=g)
    w = torch.randn(C, C, 3, 3, device="cuda", generator=g) / 17.0
    b = torch.randn(C, device="cuda", generator=g)
    res = torch.randn(N, C, H, W, device="cuda", generator=g)
    post = torch.randn(N, C, H // 2, W // 2, device="cuda", generator=g)
    def f(t):
        return 3.0 * (F.leaky_relu(F.conv2d(x.to(t), w.to(t), b.to(t), padding=1) + res.to(t), 0.2)
                      + F.interpolate(post.to(t), scale_factor=2, mode="nearest"))
    ref = f(torch.float64)
    ref32 = (f(torch.float32).double() - ref).abs().max().item()
    out = ops.conv3x3_tc_fwd(ops.nchw_to_c4(x), ops.conv_tc_pack_weights(w, 32), b, C, 32, res=ops.nchw_to_c4(res),
                             post=ops.nchw_to_c4(post), post_shift=1, act="lrelu", out_scale=3.0)
    err = (ops.c4_to_nchw(out).double() - ref).abs().max().item()
    assert err <= max(4.0 * ref32, 1e-5), (err, ref32)
