"""GPU parity of the second-generation eval conv (tcgen05 kind::f16 over PRE-SPLIT fp16-pair activations, csrc/conv_p16.cu)
and of the P16-layout helpers, through the C ABI, against torch fp64 / fp32 (cuDNN with TF32 disabled)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _no_tf32():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def _acts(shape, g, scale=20.0):
    """RelightNet-like magnitudes (|x| up to ~25) with many small values (fp16-subnormal corrections)."""
    return torch.randn(shape, device="cuda", generator=g) * torch.rand(shape, device="cuda", generator=g) ** 4 * scale


def test_p16_roundtrip_pool_upsample():
    from geomconsistentfr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    x = _acts((2, 155, 16, 16), g)
    p = ops.nchw_to_p16(x)
    assert p.data.shape == (2, 20, 2, 16, 16, 8) and p.data.dtype == torch.float16
    assert torch.equal(p.data[:, 19, :, :, :, 3:], torch.zeros(2, 2, 16, 16, 5, device="cuda", dtype=torch.float16))   # channel padding is zero
    back = ops.p16_to_nchw(p)
    assert float((back - x).abs().max()) <= 2.0 ** -21 * float(x.abs().max())          # ~22 bits survive the split
    assert float(((back - x).abs() - 2.0 ** -21 * x.abs()).max()) <= 4e-9               # + the fp16-subnormal floor of the lo part (2^-25 / 16 ... 2^-24 / 16)
    assert torch.equal(ops.p16_to_nchw(ops.nchw_to_p16(back)), back)                    # idempotent on representable values
    assert torch.equal(ops.p16_to_nchw(p, C=27, c_first=128), back[:, 128:155])
    pooled = ops.maxpool2_p16_fwd(p)
    assert torch.equal(ops.p16_to_nchw(pooled), F.max_pool2d(back, 2))
    up = ops.upsample2_p16_fwd(pooled)
    assert torch.equal(ops.p16_to_nchw(up), F.interpolate(F.max_pool2d(back, 2), scale_factor=2, mode="nearest"))


@pytest.mark.parametrize("N,Cin,Cout,H,W,cfg", [
    (2, 16, 16, 32, 32, (16, 1, 2)),
    (1, 16, 16, 256, 256, (16, 2, 2)),
    (2, 16, 16, 48, 40, (16, 2, 2)),            # ragged: tiles hang over the right / bottom edge
    (2, 32, 16, 64, 64, (16, 2, 4)),
    (2, 32, 16, 64, 64, (16, 1, 4)),
    (2, 32, 32, 64, 64, (32, 1, 4)),
    (2, 16, 32, 64, 64, (32, 2, 2)),
    (1, 64, 64, 32, 32, (32, 2, 4)),
    (2, 64, 155, 16, 16, (32, 1, 4)),
    (1, 155, 155, 16, 16, (32, 1, 4)),
    (1, 155, 155, 16, 16, (32, 1, 2)),
    (2, 128, 64, 16, 16, (32, 1, 4)),
    (8, 16, 16, 128, 128, (16, 2, 2)),          # persistent CTAs walk several tiles
])
def test_conv3x3_p16_vs_torch(N, Cin, Cout, H, W, cfg):
    from geomconsistentfr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(N * 1000 + Cin + Cout + H + cfg[1])
    x = _acts((N, Cin, H, W), g)
    w = torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / (3.0 * Cin ** 0.5)
    b = torch.randn(Cout, device="cuda", generator=g)
    xp = ops.nchw_to_p16(x)
    xq = ops.p16_to_nchw(xp)                                   # the values the kernel really sees
    ref = F.conv2d(xq.double(), w.double(), b.double(), padding=1)
    wp, w_scale = ops.conv_p16_pack_weights(w, cfg[0], cfg[2])
    flags = torch.zeros(1, dtype=torch.int32, device="cuda")
    out = ops.conv3x3_p16_fwd(xp, wp, b, Cout, cfg, w_scale, act=None, flags=flags)
    got = ops.p16_to_nchw(out)
    err = (got.double() - ref).abs().max().item()
    ref32 = (F.conv2d(xq, w, b, padding=1).double() - ref).abs().max().item()
    assert err <= max(4.0 * ref32, 2e-6 * float(ref.abs().max())), (err, ref32)
    assert int(flags.item()) == 0
    if Cout % 8:                                               # padding channels of the last chunk are exact zeros
        assert float(out.data[:, -1, :, :, :, Cout % 8:].abs().max()) == 0.0


def test_conv3x3_p16_fused_block_epilogue():
    """One launch computes a residual block's first conv (LeakyReLU) and its shortcut conv (no activation) with concatenated
    output channels; the second conv reads the leading channels as its input and the trailing ones as its residual
    operand, in place, and adds the x2-upsampled `post` operand after the activation (TRAIN:203-223, 235-246)."""
    from geomconsistentfr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(7)
    N, Cin, C, H, W = 2, 32, 64, 32, 32
    x = _acts((N, Cin, H, W), g, 5.0)
    w1 = torch.randn(C, Cin, 3, 3, device="cuda", generator=g) / 17.0
    wsc = torch.randn(C, Cin, 3, 3, device="cuda", generator=g) / 17.0
    w2 = torch.randn(C, C, 3, 3, device="cuda", generator=g) / 24.0
    b1, bsc, b2 = (torch.randn(C, device="cuda", generator=g) for _ in range(3))
    post = _acts((N, C, H // 2, W // 2), g, 5.0)
    xp, pp = ops.nchw_to_p16(x), ops.nchw_to_p16(post)
    xq, pq = ops.p16_to_nchw(xp), ops.p16_to_nchw(pp)
    cfg = (32, 1, 4)
    wcat, s_cat = ops.conv_p16_pack_weights(torch.cat([w1, wsc]), cfg[0], cfg[2])
    both = ops.conv3x3_p16_fwd(xp, wcat, torch.cat([b1, bsc]), 2 * C, cfg, s_cat, act="lrelu", act_channels=C)
    a_ref = F.leaky_relu(F.conv2d(xq.double(), w1.double(), b1.double(), padding=1), 0.2)
    sc_ref = F.conv2d(xq.double(), wsc.double(), bsc.double(), padding=1)
    got = ops.p16_to_nchw(both).double()
    assert float((got[:, :C] - a_ref).abs().max()) <= 2e-6 * float(a_ref.abs().max())
    assert float((got[:, C:] - sc_ref).abs().max()) <= 2e-6 * float(sc_ref.abs().max())
    w2p, s2 = ops.conv_p16_pack_weights(w2, cfg[0], cfg[2])
    y = ops.conv3x3_p16_fwd(both, w2p, b2, C, cfg, s2, cin=C, res=both, res_c=C, post=pp, post_shift=1, act="lrelu", out_scale=3.0)
    a_q, sc_q = ops.p16_to_nchw(both, C=C).double(), ops.p16_to_nchw(both, C=C, c_first=C).double()
    ref = 3.0 * (F.leaky_relu(F.conv2d(a_q, w2.double(), b2.double(), padding=1) + sc_q, 0.2)
                 + F.interpolate(pq.double(), scale_factor=2, mode="nearest"))
    assert float((ops.p16_to_nchw(y).double() - ref).abs().max()) <= 3e-6 * float(ref.abs().max())


def test_conv3x3_p16_overflow_flag():
    from geomconsistentfr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(9)
    x = torch.rand(1, 16, 32, 32, device="cuda", generator=g) * 20.0
    w = torch.full((16, 16, 3, 3), 4.0, device="cuda")
    b = torch.zeros(16, device="cuda")
    wp, s = ops.conv_p16_pack_weights(w, 16, 2)
    flags = torch.zeros(1, dtype=torch.int32, device="cuda")
    ops.conv3x3_p16_fwd(ops.nchw_to_p16(x), wp, b, 16, (16, 1, 2), s, act=None, flags=flags)      # outputs ~ 16*9*4*10 = 5760 > 4094
    assert int(flags.item()) == 1
    flags.zero_()
    ops.conv3x3_p16_fwd(ops.nchw_to_p16(x * 0.1), wp, b, 16, (16, 1, 2), s, act=None, flags=flags)
    assert int(flags.item()) == 0


def test_stem_head_light_p16_equal_their_c4_forms():
    from geomconsistentfr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(3)
    img = torch.rand(2, 64, 96, 3, device="cuda", generator=g)
    w = torch.randn(16, 3, 5, 5, device="cuda", generator=g) * 0.1
    b = torch.randn(16, device="cuda", generator=g)
    c = lambda t: t.cpu().contiguous()
    o4, p4 = ops.stem_conv_fwd(img, c(w), c(b))
    o16, p16 = ops.stem_conv_p16_fwd(img, c(w), c(b))
    assert torch.equal(ops.p16_to_nchw(o16), ops.p16_to_nchw(ops.nchw_to_p16(ops.c4_to_nchw(o4))))
    assert torch.equal(ops.p16_to_nchw(p16), ops.p16_to_nchw(ops.nchw_to_p16(ops.c4_to_nchw(p4))))
    x = _acts((2, 16, 32, 64), g, 3.0)
    xq = ops.p16_to_nchw(ops.nchw_to_p16(x))
    w2, w3 = (torch.randn(16, 16, device="cuda", generator=g) * 0.3 for _ in range(2))
    b2, b3 = (torch.randn(16, device="cuda", generator=g) * 0.1 for _ in range(2))
    wo, bo = torch.randn(3, 16, device="cuda", generator=g) * 0.3, torch.randn(3, device="cuda", generator=g) * 0.1
    h4 = ops.head_1x1_fwd(ops.nchw_to_c4(xq), c(w2), c(b2), c(w3), c(b3), c(wo), c(bo), act="sigmoid")
    h16 = ops.head_1x1_p16_fwd(ops.nchw_to_p16(xq), c(w2), c(b2), c(w3), c(b3), c(wo), c(bo), act="sigmoid")
    assert torch.equal(h4, h16)
    f = _acts((3, 155, 16, 16), g, 3.0)
    fq = ops.p16_to_nchw(ops.nchw_to_p16(f))
    w1, b1 = torch.randn(128, 27, device="cuda", generator=g) * 0.2, torch.randn(128, device="cuda", generator=g) * 0.1
    wl, bl = torch.randn(4, 128, device="cuda", generator=g) * 0.2, torch.randn(4, device="cuda", generator=g) * 0.1
    l4 = ops.light_head_c4_fwd(ops.nchw_to_c4(fq), 128, w1, b1, wl, bl)
    l16 = ops.light_head_p16_fwd(ops.nchw_to_p16(fq), 128, w1, b1, wl, bl)
    assert float((l4 - l16).abs().max()) <= 1e-6


@pytest.mark.parametrize("N,H,W,MH,n_out,act,scale", [(2, 64, 96, 2, 3, "sigmoid", 1.0), (1, 256, 256, 2, 1, None, 100.0),
                                                       (3, 40, 24, 1, 1, None, 100.0), (8, 128, 128, 1, 3, "sigmoid", 1.0)])
def test_conv_with_fused_1x1_tail_equals_the_two_launches(N, H, W, MH, n_out, act, scale):
    """gfr_conv3x3_p16_head_fwd (the decoder's 1x1 tail in the epilogue of its last 3x3 layer) against conv3x3_p16_fwd +
    head_1x1_p16_fwd: the same arithmetic except that the 16-channel activation is not rounded to the fp16 pair (22 bits) in
    between — tolerance 2e-6 of the output range — and against torch fp64."""
    from geomconsistentfr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(N + H + n_out)
    x = _acts((N, 16, H, W), g, 4.0)
    w = torch.randn(16, 16, 3, 3, device="cuda", generator=g) / 12.0
    b = torch.randn(16, device="cuda", generator=g) * 0.1
    w2, w3 = (torch.randn(16, 16, device="cuda", generator=g) * 0.3 for _ in range(2))
    b2, b3 = (torch.randn(16, device="cuda", generator=g) * 0.1 for _ in range(2))
    wo, bo = torch.randn(n_out, 16, device="cuda", generator=g) * 0.3, torch.randn(n_out, device="cuda", generator=g) * 0.1
    c = lambda t: t.cpu().contiguous()
    xp = ops.nchw_to_p16(x)
    wp, s = ops.conv_p16_pack_weights(w, 16, 2)
    two = ops.head_1x1_p16_fwd(ops.conv3x3_p16_fwd(xp, wp, b, 16, (16, MH, 2), s), c(w2), c(b2), c(w3), c(b3), c(wo), c(bo), act=act, out_scale=scale)
    one = ops.conv3x3_p16_head_fwd(xp, wp, b, MH, s, c(w2), c(b2), c(w3), c(b3), c(wo), c(bo), act=act, out_scale=scale)
    assert one.shape == two.shape == (N, n_out, H, W)
    tol = 2e-6 * max(1.0, float(two.abs().max()))
    assert float((one - two).abs().max()) <= tol, float((one - two).abs().max())
    xq = ops.p16_to_nchw(xp).double()
    lr = lambda v: F.leaky_relu(v, 0.2)
    h = lr(F.conv2d(xq, w.double(), b.double(), padding=1))
    h = lr(F.conv2d(h, w2.double()[:, :, None, None], b2.double()))
    h = lr(F.conv2d(h, w3.double()[:, :, None, None], b3.double()))
    ref = F.conv2d(h, wo.double()[:, :, None, None], bo.double())
    ref = (torch.sigmoid(ref) if act == "sigmoid" else ref) * scale
    assert float((one.double() - ref).abs().max()) <= 1e-5 * max(1.0, float(ref.abs().max()))


def test_stem_on_tensor_cores_and_fused_pool():
    """The 5x5 stem as horizontal-tap unroll + 5 vertical taps on tcgen05 (geometry 1) against torch fp64, and the 2x2 max pool
    that the epilogue writes alongside (stem: no residual; a 3x3 layer with a residual operand and LeakyReLU)."""
    from geomconsistentfr_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(11)
    img = torch.rand(2, 64, 96, 3, device="cuda", generator=g)
    w = torch.randn(16, 3, 5, 5, device="cuda", generator=g) * 0.1
    b = torch.randn(16, device="cuda", generator=g) * 0.1
    w5 = w.new_zeros((16, 16, 5))
    w5[:, :15, :] = w.permute(0, 3, 1, 2).reshape(16, 15, 5)
    wp, ws = ops.conv_p16_pack_weights_taps(w5, 16, 2)
    u = ops.stem_unroll_p16(img)
    un = ops.p16_to_nchw(u)
    x = img.permute(0, 3, 1, 2)
    for kx in range(5):                                        # the unroll itself: channel kx*3 + c = img shifted by kx - 2
        sh = F.pad(x, (2, 2, 0, 0))[:, :, :, kx:kx + 96]
        assert float((un[:, kx * 3:kx * 3 + 3] - sh).abs().max()) <= 2.0 ** -21
    assert float(un[:, 15].abs().max()) == 0.0
    full, pooled = ops.conv3x3_p16_fwd(u, wp, b, 16, (16, 2, 2), ws, pool=True, geometry=1)
    ref = F.leaky_relu(F.conv2d(x.double(), w.double(), b.double(), padding=2), 0.2)
    got = ops.p16_to_nchw(full)
    assert float((got.double() - ref).abs().max()) <= 3e-6 * float(ref.abs().max())
    assert torch.equal(ops.p16_to_nchw(pooled), F.max_pool2d(got, 2))
    # fused pool behind residual + activation, all three tile shapes
    for (N, C, S, cfg) in ((2, 16, 64, (16, 2, 2)), (2, 32, 32, (32, 1, 4)), (1, 16, 32, (16, 1, 2))):
        xx = _acts((N, C, S, S), g, 4.0)
        ww = torch.randn(C, C, 3, 3, device="cuda", generator=g) / (3.0 * C ** 0.5)
        bb = torch.randn(C, device="cuda", generator=g)
        rr = _acts((N, C, S, S), g, 4.0)
        wq, sq = ops.conv_p16_pack_weights(ww, cfg[0], cfg[2])
        y, yp = ops.conv3x3_p16_fwd(ops.nchw_to_p16(xx), wq, bb, C, cfg, sq, res=ops.nchw_to_p16(rr), pool=True)
        y0 = ops.conv3x3_p16_fwd(ops.nchw_to_p16(xx), wq, bb, C, cfg, sq, res=ops.nchw_to_p16(rr))
        assert torch.equal(ops.p16_to_nchw(y), ops.p16_to_nchw(y0))
        assert torch.equal(ops.p16_to_nchw(yp), F.max_pool2d(ops.p16_to_nchw(y), 2))
