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


def test_stem_conv_vs_torch():
    """conv_c1_og 5x5 3->16 + folded BN + LeakyReLU + fused 2x2 max pool on the NHWC image (TRAIN:197-201); exact fp32."""
    from geomconsistentfr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    img = torch.rand(2, 64, 96, 3, device="cuda", generator=g)
    w = torch.randn(16, 3, 5, 5, device="cuda", generator=g) * 0.1
    b = torch.randn(16, device="cuda", generator=g)
    ref = F.leaky_relu(F.conv2d(img.permute(0, 3, 1, 2), w, b, padding=2), 0.2)
    out, pooled = ops.stem_conv_fwd(img, w.cpu().contiguous(), b.cpu().contiguous())
    assert (ops.c4_to_nchw(out) - ref).abs().max().item() <= 5e-6
    assert (ops.c4_to_nchw(pooled) - F.max_pool2d(ref, 2)).abs().max().item() <= 5e-6


@pytest.mark.parametrize("n_out,act,scale", [(3, "sigmoid", 1.0), (1, None, 100.0)])
def test_head_1x1_vs_torch(n_out, act, scale):
    """c2_2 -> c2_3 -> c2_o 1x1 chain fused per pixel (TRAIN:285-290, 345-350); exact fp32."""
    from geomconsistentfr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(4)
    x = torch.randn(2, 16, 32, 64, device="cuda", generator=g)
    w2, w3 = (torch.randn(16, 16, device="cuda", generator=g) * 0.3 for _ in range(2))
    b2, b3 = (torch.randn(16, device="cuda", generator=g) * 0.1 for _ in range(2))
    wo = torch.randn(n_out, 16, device="cuda", generator=g) * 0.3
    bo = torch.randn(n_out, device="cuda", generator=g) * 0.1
    h = F.leaky_relu(F.conv2d(x, w2.view(16, 16, 1, 1), b2), 0.2)
    h = F.leaky_relu(F.conv2d(h, w3.view(16, 16, 1, 1), b3), 0.2)
    ref = F.conv2d(h, wo.view(n_out, 16, 1, 1), bo)
    ref = (torch.sigmoid(ref) if act == "sigmoid" else ref) * scale
    c = lambda t: t.cpu().contiguous()
    out = ops.head_1x1_fwd(ops.nchw_to_c4(x), c(w2), c(b2), c(w3), c(b3), c(wo), c(bo), act=act, out_scale=scale)
    assert (out - ref).abs().max().item() <= 2e-6 * scale + 1e-6


def test_conv3x3_tc_channel_slice():
    """cin < C reads the leading channels of a wider C4 tensor in place (TRAIN:225 feature split)."""
    from geomconsistentfr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.randn(2, 155, 16, 16, device="cuda", generator=g)
    w = torch.randn(64, 128, 3, 3, device="cuda", generator=g) / 34.0
    b = torch.zeros(64, device="cuda")
    ref = F.conv2d(x[:, :128], w, b, padding=1)
    out = ops.conv3x3_tc_fwd(ops.nchw_to_c4(x), ops.conv_tc_pack_weights(w, 32), b, 64, 32, act=None, cin=128)
    assert (ops.c4_to_nchw(out) - ref).abs().max().item() <= 1e-5


@pytest.mark.parametrize("N,Cin,Cout,S", [(2, 16, 16, 64), (2, 32, 64, 32), (1, 155, 155, 16), (2, 128, 64, 16)])
def test_conv3x3_tc_fp16_pair_split_vs_torch(N, Cin, Cout, S):
    """precision=2 (kind::f16, x = h1 + h2 and w = w1 + w2 as scaled fp16 pairs): same error class as 3xTF32, on
    activations of RelightNet-like magnitude (|x| up to ~25) including small values (fp16-subnormal corrections)."""
    from geomconsistentfr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(Cin * 7 + Cout)
    x = torch.randn(N, Cin, S, S, device="cuda", generator=g) * torch.rand(N, Cin, S, S, device="cuda", generator=g) ** 4 * 20.0
    w = torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / (3.0 * Cin ** 0.5)
    b = torch.randn(Cout, device="cuda", generator=g)
    ref = F.conv2d(x.double(), w.double(), b.double(), padding=1)
    wp, w_scale, NT = ops.conv_tc_pack_weights_f16(w)
    out = ops.conv3x3_tc_fwd(ops.nchw_to_c4(x), wp, b, Cout, NT, act=None, precision=2, w_scale=w_scale)
    err = (ops.c4_to_nchw(out).double() - ref).abs().max().item()
    ref32 = (F.conv2d(x, w, b, padding=1).double() - ref).abs().max().item()
    assert err <= max(4.0 * ref32, 2e-6 * float(ref.abs().max())), (err, ref32)
