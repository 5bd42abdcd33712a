"""GPU parity of the individual kernels behind the C ABI against plain torch fp32 (floating-point kernels keep a
torch fp32 reference; tolerances are stated per test and reflect the operand precision: bf16 2^-9, tf32 2^-11)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "-m gpu tests need a CUDA device"
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch.device("cuda:0")


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def _dt(name):
    from nlc_b200._lib import NLC_BF16, NLC_F16, NLC_F32
    return {"bf16": NLC_BF16, "fp16": NLC_F16}.get(name, NLC_F32)


def _rnd(x, dt):
    from nlc_b200 import ops
    from nlc_b200._lib import NLC_BF16, NLC_F16
    if dt in (NLC_BF16, NLC_F16):
        return x.to(ops.OP_DTYPES[dt]).float()
    return ops.round_tf32_(x.clone())


CONV_CASES = [
    # B, H, W, Cin, Cout, stride, pad(l,r,t,b), k
    (4, 16, 16, 128, 128, 1, (1, 1, 1, 1), 3),
    (5, 4, 4, 256, 512, 1, (1, 1, 1, 1), 3),       # ragged batch: 5 images of 16 pixels in 128-row tiles
    (3, 32, 32, 256, 768, 1, (0, 0, 0, 0), 1),     # qkv-shaped 1x1
    (4, 32, 32, 128, 128, 2, (0, 1, 0, 1), 3),     # DDIM Downsample: pad right/bottom, stride 2
    (6, 8, 8, 256, 256, 2, (1, 1, 1, 1), 3),       # ADM Downsample: pad 1, stride 2
    (1, 256, 256, 64, 64, 1, (1, 1, 1, 1), 3),     # widest tile row
    (2, 2, 2, 512, 512, 1, (1, 1, 1, 1), 3),       # sigma-model 2x2 level
    (130, 1, 1, 256, 128, 1, (0, 0, 0, 0), 1),     # 1x1 spatial, batch > one tile
]


@pytest.mark.parametrize("prec", ["bf16", "tf32", "fp16"])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_tc_matches_torch(dev, prec, case):
    """Operands are pre-rounded to the operand dtype, so the only difference left is fp32 summation order:
    tolerance 5e-5 of the output's max magnitude."""
    from nlc_b200 import ops
    B, H, W, Cin, Cout, stride, pad, k = case
    dt = _dt(prec)
    g = torch.Generator().manual_seed(1)
    x = _rnd(torch.randn(B, Cin, H, W, generator=g).to(dev), dt)
    w = _rnd((torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).to(dev), dt)
    b = torch.randn(Cout, generator=g).to(dev)
    ref = F.conv2d(F.pad(x, pad), w, b, stride=stride)
    Ho, Wo = ref.shape[2:]
    rowvec = torch.randn(B, Cout, generator=g).to(dev)
    resid = torch.randn(B, Ho, Wo, Cout, generator=g).to(dev)
    ref = (ref + rowvec[:, :, None, None] + resid.permute(0, 3, 1, 2)) * 0.5
    tdt = ops.OP_DTYPES[dt]
    xa = ops.Act(x.permute(0, 2, 3, 1).contiguous().to(tdt))
    segs = [(0, kh - pad[2], kw - pad[0], 0, Cin) for kh in range(k) for kw in range(k)]
    o32 = ops.Act(torch.full((B, Ho, Wo, Cout), float("nan"), device=dev))
    oop = ops.Act(torch.zeros(B, Ho, Wo, Cout, device=dev, dtype=tdt))
    ops.conv_tc([xa], segs, ops.pack_conv_weight(w, dt), Cout, B, Ho, Wo, dt, stride=stride, bias=b, rowvec=rowvec,
                resid=ops.Act(resid), out_scale=0.5, out_f32=o32, out_op=oop)
    torch.cuda.synchronize()
    assert _rel(o32.t.permute(0, 3, 1, 2), ref) < 5e-5
    assert _rel(oop.t.float().permute(0, 3, 1, 2), ref) < (8e-3 if prec == "bf16" else 1e-3)


SLAB_CASES = [
    # B, H, W, Cin, Cout, extra 1x1 channels
    (3, 64, 64, 128, 128, 0),    # the c2 / c3 shape
    (2, 64, 64, 256, 128, 0),    # concat input, 12 slabs per unit
    (3, 16, 16, 128, 128, 0),    # one super tile per image, odd count: the pair's second CTA runs past the end
    (2, 32, 32, 128, 256, 0),    # two n tiles
    (2, 64, 64, 128, 128, 192),  # fused 1x1 shortcut over a second tensor (ResBlock with nin_shortcut)
    (1, 32, 64, 64, 128, 64),    # H != W
]


@pytest.mark.parametrize("prec", ["bf16", "tf32", "fp16"])
@pytest.mark.parametrize("case", SLAB_CASES)
def test_conv_slab_matches_torch(dev, prec, case):
    """The halo-slab kernel (conv_slab.cu: one slab per channel chunk and horizontal tap feeds the three vertical taps of two
    M tiles) against torch, forced on for every eligible layer, with the full epilogue (bias, per-sample row, residual, scale,
    GroupNorm partials, fp32 + operand outputs); and against the tap-per-tile kernel on the same inputs (same products,
    different summation order)."""
    from nlc_b200 import _lib, ops
    B, H, W, Cin, Cout, Cx = case
    dt = _dt(prec)
    tdt = ops.OP_DTYPES[dt]
    g = torch.Generator().manual_seed(23)
    x = _rnd(torch.randn(B, Cin, H, W, generator=g).to(dev), dt)
    w = _rnd((torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5).to(dev), dt)
    b = torch.randn(Cout, generator=g).to(dev)
    rowvec = torch.randn(B, Cout, generator=g).to(dev)
    resid = torch.randn(B, H, W, Cout, generator=g).to(dev)
    ref = F.conv2d(x, w, b, padding=1)
    srcs = [ops.Act(x.permute(0, 2, 3, 1).contiguous().to(tdt))]
    segs = ops.taps3x3(0, 0, Cin)
    extra = None
    if Cx:
        wide = _rnd(torch.randn(B, Cx + 64, H, W, generator=g).to(dev), dt)  # the shortcut reads channels [64, 64 + Cx)
        wx = _rnd((torch.randn(Cout, Cx, 1, 1, generator=g) / Cx ** 0.5).to(dev), dt)
        ref = ref + F.conv2d(wide[:, 64:], wx)
        srcs.append(ops.Act(wide.permute(0, 2, 3, 1).contiguous().to(tdt)))
        segs = segs + [(1, 0, 0, 64, Cx)]
        extra = wx
    ref = (ref + rowvec[:, :, None, None] + resid.permute(0, 3, 1, 2)) * 0.5
    wp = ops.pack_conv_weight(w, dt, extra)
    outs = {}
    ctx = _lib.ctx(0)
    try:
        for mode in (2, 0):
            _lib.check(_lib.lib().nlc_ctx_set(ctx, b"slab", mode))
            st = ops.GnStats(torch.zeros(B * H * W // 32, Cout // 4, 2, device=dev))
            o32 = ops.Act(torch.full((B, H, W, Cout), float("nan"), device=dev), 0, Cout, st)
            oop = ops.Act(torch.zeros(B, H, W, Cout, device=dev, dtype=tdt))
            ops.conv_tc(srcs, segs, wp, Cout, B, H, W, dt, bias=b, rowvec=rowvec, resid=ops.Act(resid), out_scale=0.5,
                        out_f32=o32, out_op=oop, stats=True)
            torch.cuda.synchronize()
            outs[mode] = (o32.t.clone(), oop.t.float().clone(), st.t.clone())
    finally:
        _lib.check(_lib.lib().nlc_ctx_set(ctx, b"slab", 1))
    o32, oop, stats = outs[2]
    assert _rel(o32.permute(0, 3, 1, 2), ref) < 5e-5
    assert _rel(oop.permute(0, 3, 1, 2), ref) < (8e-3 if prec == "bf16" else 1e-3)
    assert _rel(o32, outs[0][0]) < 2e-5
    # GroupNorm partials: (mean, M2) per 32 pixels x 4 channels
    blocks = o32.reshape(B * H * W // 32, 32, Cout // 4, 4).permute(0, 2, 1, 3).reshape(-1, Cout // 4, 128).double()
    assert (stats[:, :, 0].double() - blocks.mean(2)).abs().max() < 1e-4 * blocks.abs().max()
    m2 = ((blocks - blocks.mean(2, keepdim=True)) ** 2).sum(2)
    assert ((stats[:, :, 1].double() - m2).abs() / m2.clamp_min(1e-3)).max() < 1e-3


@pytest.mark.parametrize("prec", ["bf16", "tf32"])
def test_conv_tc_fused_shortcut_and_channel_slices(dev, prec):
    """3x3 over one source + 1x1 shortcut over a channel slice of a wider buffer, one accumulator."""
    from nlc_b200 import ops
    dt = _dt(prec)
    B, H, C0, C1, Cout = 4, 16, 128, 256, 128
    g = torch.Generator().manual_seed(2)
    x0 = _rnd(torch.randn(B, C0, H, H, generator=g).to(dev), dt)
    x1 = _rnd(torch.randn(B, C1, H, H, generator=g).to(dev), dt)
    w0 = _rnd((torch.randn(Cout, C0, 3, 3, generator=g) / (C0 * 9) ** 0.5).to(dev), dt)
    w1 = _rnd((torch.randn(Cout, C1, 1, 1, generator=g) / C1 ** 0.5).to(dev), dt)
    ref = F.conv2d(x0, w0, padding=1) + F.conv2d(x1, w1)
    tdt = ops.OP_DTYPES[dt]
    buf = torch.zeros(B, H, H, C1 + 64, device=dev, dtype=tdt)
    buf[..., 64:] = x1.permute(0, 2, 3, 1).to(tdt)
    out_wide = torch.zeros(B, H, H, Cout + 128, device=dev)
    out = ops.Act(out_wide, 128, Cout)
    ops.conv_tc([ops.Act(x0.permute(0, 2, 3, 1).contiguous().to(tdt)), ops.Act(buf, 64, C1)],
                ops.taps3x3(0, 0, C0) + [(1, 0, 0, 0, C1)], ops.pack_conv_weight(w0, dt, extra=w1), Cout, B, H, H, dt,
                out_f32=out)
    torch.cuda.synchronize()
    assert _rel(out.dense().permute(0, 3, 1, 2), ref) < 5e-5
    assert out_wide[..., :128].abs().max() == 0  # the neighbouring slice is untouched


def test_conv_tc_rejects_bad_shapes(dev):
    from nlc_b200 import ops
    from nlc_b200._lib import NLC_BF16, NlcError
    x = ops.Act(torch.zeros(1, 6, 6, 64, device=dev, dtype=torch.bfloat16))
    w = torch.zeros(64, 64, device=dev, dtype=torch.bfloat16)
    with pytest.raises(NlcError):  # 6x6 is not a power-of-two extent
        ops.conv_tc([x], [(0, 0, 0, 0, 64)], w, 64, 1, 6, 6, NLC_BF16, out_f32=ops.Act(torch.zeros(1, 6, 6, 64, device=dev)))
    x = ops.Act(torch.zeros(1, 8, 8, 48, device=dev, dtype=torch.bfloat16))
    with pytest.raises(NlcError):  # 48 channels do not fill a 128-byte K chunk
        ops.conv_tc([x], [(0, 0, 0, 0, 48)], w, 64, 1, 8, 8, NLC_BF16, out_f32=ops.Act(torch.zeros(1, 8, 8, 64, device=dev)))


@pytest.mark.parametrize("prec,tol", [("bf16", 6e-3), ("tf32", 8e-4), ("fp16", 8e-4)])
@pytest.mark.parametrize("shape", [(3, 8, 8, 256), (2, 64, 64, 384), (5, 2, 2, 512), (2, 16, 16, 1024)])
@pytest.mark.parametrize("silu,ss", [(True, False), (False, True)])
def test_groupnorm(dev, prec, tol, shape, silu, ss):
    """tolerance = one rounding of the output to the operand dtype (bf16 2^-8 / tf32 2^-11 of max |y|)."""
    from nlc_b200 import ops
    B, H, W, C = shape
    dt = _dt(prec)
    g = torch.Generator().manual_seed(3)
    x = torch.randn(B, C, H, W, generator=g).to(dev) * 2 + 0.5
    gam, bet = torch.randn(C, generator=g).to(dev), torch.randn(C, generator=g).to(dev)
    ref = F.group_norm(x, 32, gam, bet, eps=1e-6)
    sc = sh = None
    if ss:
        sc, sh = torch.randn(B, C, generator=g).to(dev), torch.randn(B, C, generator=g).to(dev)
        ref = ref * (1 + sc[:, :, None, None]) + sh[:, :, None, None]
    if silu:
        ref = F.silu(ref)
    y = ops.Act(torch.zeros(B, H, W, C, device=dev, dtype=ops.OP_DTYPES[dt]))
    ws = torch.zeros(ops.groupnorm_ws(B, H * W, C, 32), device=dev)
    ops.groupnorm(ops.Act(x.permute(0, 2, 3, 1).contiguous()), 32, 1e-6, gam, bet, y, dt, ws, silu=silu, scale=sc, shift=sh)
    assert _rel(y.t.float().permute(0, 3, 1, 2), ref) < tol


def test_groupnorm_large_mean_is_stable(dev):
    """Welford partials: a mean 1000x the standard deviation must not cancel."""
    from nlc_b200 import ops
    from nlc_b200._lib import NLC_F32
    B, H, W, C = 2, 32, 32, 128
    g = torch.Generator().manual_seed(4)
    x = (torch.randn(B, C, H, W, generator=g) * 0.01 + 10.0).to(dev)
    ref = F.group_norm(x.double(), 32, eps=1e-6).float()
    y = ops.Act(torch.zeros(B, H, W, C, device=dev))
    ws = torch.zeros(ops.groupnorm_ws(B, H * W, C, 32), device=dev)
    ops.groupnorm(ops.Act(x.permute(0, 2, 3, 1).contiguous()), 32, 1e-6, None, None, y, NLC_F32, ws, silu=False)
    assert (y.t.permute(0, 3, 1, 2) - ref).abs().max() < 2e-2  # fp32 input quantisation of x itself is ~1e-4 sigma


ATTN_CASES = [(3, 16, 1, 512, False), (2, 64, 4, 64, True), (2, 64, 1, 256, False), (3, 256, 1, 256, False),
              (2, 1024, 4, 64, False), (2, 256, 4, 64, True),
              # the fused tcgen05 kernel (bf16, 64-channel heads, T % 128 == 0): one key block, odd batch, more tiles
              # than SMs (several tiles per persistent CTA), both qkv orders
              (3, 128, 2, 64, True), (1, 512, 8, 64, False), (5, 256, 16, 64, True), (3, 1024, 8, 64, True),
              # head dimension 256 (unet_ddim / SongUNet single-head blocks): one CTA per SM, four 64-channel chunks
              (160, 256, 1, 256, False), (2, 1024, 1, 256, False), (3, 128, 2, 256, True)]


@pytest.mark.parametrize("prec,tol", [("bf16", 8e-3), ("tf32", 1e-3), ("fp16", 1e-3)])
@pytest.mark.parametrize("case", ATTN_CASES)
def test_attention(dev, prec, tol, case):
    from nlc_b200 import ops
    B, T, heads, dh, legacy = case
    dt = _dt(prec)
    C = heads * dh
    tdt = ops.OP_DTYPES[dt]
    g = torch.Generator().manual_seed(5)
    qkv = _rnd(torch.randn(B, T, 3 * C, generator=g).to(dev), dt).to(tdt)
    f = qkv.float()
    if legacy:
        v5 = f.view(B, T, heads, 3, dh)
        q, k, v = v5[:, :, :, 0], v5[:, :, :, 1], v5[:, :, :, 2]
        offs = (0, dh, 2 * dh, 3 * dh)
    else:
        q, k, v = [f[:, :, i * C:(i + 1) * C].view(B, T, heads, dh) for i in range(3)]
        offs = (0, C, 2 * C, dh)
    scale = dh ** -0.5
    w = torch.softmax(torch.einsum("bthd,bshd->bhts", q, k) * scale, dim=-1)
    ref = torch.einsum("bhts,bshd->bthd", w, v).reshape(B, T, C)
    side = 1 << ((T.bit_length() - 1) // 2)  # H x W = T (T is a power of two, not always a square)
    out = ops.Act(torch.zeros(B, side, T // side, C, device=dev, dtype=tdt))
    ws = torch.zeros(max(ops.attention_ws(dt, B, T, heads, dh), 16), device=dev, dtype=torch.uint8)
    ops.attention(ops.Act(qkv.view(B, side, T // side, 3 * C)), dt, offs[0], offs[1], offs[2], offs[3], heads, dh, scale,
                  out, ws)
    assert _rel(out.t.float().view(B, T, C), ref) < tol


@pytest.mark.parametrize("prec,tol", [("bf16", 1.2e-2), ("fp16", 2e-3)])
@pytest.mark.parametrize("case", [(2, 1024, 4, 64), (3, 256, 1, 256), (2, 64, 2, 64), (1, 512, 2, 256)])
def test_attention_online_softmax_rescales(dev, prec, tol, case):
    """One-pass fused attention (attn_fused1_kernel) on logits whose magnitude GROWS along the keys (key s scaled by a ramp
    0.2 -> 9), so that the running reference of most rows moves several times and the accumulator in TMEM is rescaled; both
    kernels (one pass / two passes) against torch's fp32 softmax on the same 16-bit inputs."""
    from nlc_b200 import _lib, ops
    B, T, heads, dh = case
    dt = _dt(prec)
    C = heads * dh
    tdt = ops.OP_DTYPES[dt]
    g = torch.Generator().manual_seed(11)
    q = torch.randn(B, T, heads, dh, generator=g)
    k = torch.randn(B, T, heads, dh, generator=g) * torch.linspace(0.2, 9.0, T).view(1, T, 1, 1)
    v = torch.randn(B, T, heads, dh, generator=g)
    qkv = torch.cat([q.reshape(B, T, C), k.reshape(B, T, C), v.reshape(B, T, C)], dim=2).to(dev).to(tdt)
    f = qkv.float()
    q, k, v = [f[:, :, i * C:(i + 1) * C].view(B, T, heads, dh) for i in range(3)]
    scale = dh ** -0.5
    logits = torch.einsum("bthd,bshd->bhts", q, k) * scale
    # the premise of the test: block maxima (64 keys) of a typical row climb by far more than the 2^8 rescale threshold
    blk = logits.view(B, heads, T, T // 64, 64).amax(-1) * 1.4426950408889634
    if T > 64:
        assert ((blk.cummax(-1).values[..., -1] - blk[..., 0]) > 16).float().mean() > 0.5
    ref = torch.einsum("bhts,bshd->bthd", torch.softmax(logits, dim=-1), v).reshape(B, T, C)
    side = 1 << ((T.bit_length() - 1) // 2)
    ws = torch.zeros(max(ops.attention_ws(dt, B, T, heads, dh), 16), device=dev, dtype=torch.uint8)
    ctx = _lib.ctx(0)
    try:
        for mode in (1, 0):
            _lib.check(_lib.lib().nlc_ctx_set(ctx, b"attn_onepass", mode))
            out = ops.Act(torch.zeros(B, side, T // side, C, device=dev, dtype=tdt))
            ops.attention(ops.Act(qkv.view(B, side, T // side, 3 * C)), dt, 0, C, 2 * C, dh, heads, dh, scale, out, ws)
            torch.cuda.synchronize()
            assert _rel(out.t.float().view(B, T, C), ref) < tol, "mode %d" % mode
    finally:
        _lib.check(_lib.lib().nlc_ctx_set(ctx, b"attn_onepass", 1))


def test_linear_and_embedding(dev):
    from nlc_b200 import ops
    g = torch.Generator().manual_seed(6)
    x = torch.randn(37, 300, generator=g).to(dev)
    W = (torch.randn(130, 300, generator=g) / 17).to(dev)
    b = torch.randn(130, generator=g).to(dev)
    y = torch.zeros(37, 130, device=dev)
    ops.linear(x, W, b, y, act_in=1, act_out=2)
    assert _rel(y, F.gelu(F.linear(F.silu(x), W, b))) < 2e-6
    # K not a multiple of 4 (scalar-load path), and a strided row view of a wider buffer (vector path, ld_x > K)
    x1, W1 = torch.randn(5, 301, generator=g).to(dev), (torch.randn(70, 301, generator=g) / 17).to(dev)
    y1 = torch.zeros(5, 70, device=dev)
    ops.linear(x1, W1, None, y1)
    assert _rel(y1, F.linear(x1, W1)) < 2e-6
    wide = torch.randn(33, 1024 + 64, generator=g).to(dev)
    W2 = (torch.randn(4100, 1024, generator=g) / 32).to(dev)
    ywide = torch.zeros(33, 4100 + 12, device=dev)
    ops.linear(wide[:, 64:], W2, None, ywide[:, 12:], act_in=1)
    assert _rel(ywide[:, 12:], F.linear(F.silu(wide[:, 64:]), W2)) < 2e-6 and ywide[:, :12].abs().max() == 0
    t = torch.tensor([0.0, 1.0, 37.0, 999.0, 1000.0], device=dev)
    freqs = torch.exp(torch.arange(64, dtype=torch.float32) * -(9.210340371976184 / 63)).to(dev)
    out = torch.zeros(5, 128, device=dev)
    ops.timestep_embedding(t, freqs, False, out)
    arg = t[:, None] * freqs[None]
    assert (out - torch.cat([arg.sin(), arg.cos()], 1)).abs().max() < 2e-6


@pytest.mark.parametrize("prec", ["bf16", "tf32", "fp16"])
def test_resample_and_boundary_convs(dev, prec):
    from nlc_b200 import ops
    dt = _dt(prec)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(2, 64, 8, 8, generator=g).to(dev)
    xa = ops.Act(x.permute(0, 2, 3, 1).contiguous())
    for mode, ref in ((0, x), (1, F.interpolate(x, scale_factor=2.0, mode="nearest")), (2, F.avg_pool2d(x, 2))):
        B, C, Ho, Wo = ref.shape
        yf = ops.Act(torch.zeros(B, Ho, Wo, C, device=dev))
        ops.resample(xa, mode, yf, None, dt)
        assert _rel(yf.t.permute(0, 3, 1, 2), ref) < 1e-6
    B, R, C = 3, 16, 128
    img = torch.randn(B, 3, R, R, generator=g).to(dev)
    sc = (torch.rand(B, generator=g) + 0.5).to(dev)
    w = (torch.randn(C, 3, 3, 3, generator=g) / 5).to(dev)
    b = torch.randn(C, generator=g).to(dev)
    yf = ops.Act(torch.zeros(B, R, R, C, device=dev))
    yo = ops.Act(torch.zeros(B, R, R, C, device=dev, dtype=ops.OP_DTYPES[dt]))
    ops.conv_in_nchw(img, sc, w, b, yf, yo, dt)
    assert _rel(yf.t.permute(0, 3, 1, 2), F.conv2d(img * sc[:, None, None, None], w, b, padding=1)) < 2e-6
    w2 = (torch.randn(3, C, 3, 3, generator=g) / 30).to(dev)
    b2 = torch.randn(3, generator=g).to(dev)
    out = torch.zeros(B, 3, R, R, device=dev)
    ops.conv_out_nchw(yo, dt, w2, b2, out)
    assert _rel(out, F.conv2d(yo.t.float().permute(0, 3, 1, 2), w2, b2, padding=1)) < 5e-6


@pytest.mark.parametrize("prec,tol", [("bf16", 6e-3), ("tf32", 8e-4)])
@pytest.mark.parametrize("shape", [(2, 16, 16, 128, 256, 128), (3, 64, 64, 128, 128, 128), (1, 256, 256, 64, 192, 64),
                                   (2, 32, 8, 64, 384, 128)])
def test_groupnorm_from_conv_epilogue_statistics(dev, prec, tol, shape):
    """The conv epilogue's per-(32 px, 4 ch) partials + nlc_groupnorm(stats=...) == F.group_norm of the conv output,
    including a GroupNorm over a *concatenation* whose two halves were written by different convolutions (group
    boundaries straddle the seam: 256+128 channels in 32 groups of 12) and a large common offset (mean >> std)."""
    from nlc_b200 import ops
    B, H, W, Cin, Cout, Cb = shape
    dt = _dt(prec)
    tdt = ops.OP_DTYPES[dt]
    g = torch.Generator().manual_seed(9)
    x = _rnd(torch.randn(B, Cin, H, W, generator=g).to(dev), dt)
    Ct = Cout + Cb  # second producer: a 1x1 conv writing the tail channels of the same buffer
    w1 = _rnd((torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5).to(dev), dt)
    w2 = _rnd((torch.randn(Cb, Cin, 1, 1, generator=g) / Cin ** 0.5).to(dev), dt)
    b1 = (torch.randn(Cout, generator=g) + 30.0).to(dev)  # large mean
    b2 = torch.randn(Cb, generator=g).to(dev)
    ref_t = torch.cat([F.conv2d(x, w1, b1, padding=1), F.conv2d(x, w2, b2)], dim=1)
    xa = ops.Act(x.permute(0, 2, 3, 1).contiguous().to(tdt))
    buf = torch.zeros(B, H, W, Ct, device=dev)
    st = ops.GnStats(torch.zeros(B * H * W // 32, Ct // 4, 2, device=dev))
    o1, o2 = ops.Act(buf, 0, Cout, st), ops.Act(buf, Cout, Cb, st)
    ops.conv_tc([xa], ops.taps3x3(0, 0, Cin), ops.pack_conv_weight(w1, dt), Cout, B, H, W, dt, bias=b1, out_f32=o1, stats=True)
    ops.conv_tc([xa], [(0, 0, 0, 0, Cin)], ops.pack_conv_weight(w2, dt), Cb, B, H, W, dt, bias=b2, out_f32=o2, stats=True)
    assert _rel(buf.permute(0, 3, 1, 2), ref_t) < 5e-5
    gam, bet = torch.randn(Ct, generator=g).to(dev), torch.randn(Ct, generator=g).to(dev)
    ref = F.silu(F.group_norm(buf.permute(0, 3, 1, 2), 32, gam, bet, eps=1e-5))
    y = ops.Act(torch.zeros(B, H, W, Ct, device=dev, dtype=tdt))
    ws = torch.zeros(ops.groupnorm_ws(B, H * W, Ct, 32), device=dev)
    ops.groupnorm(ops.Act(buf, 0, Ct, st), 32, 1e-5, gam, bet, y, dt, ws, silu=True, use_stats=True)
    assert _rel(y.t.float().permute(0, 3, 1, 2), ref) < tol
    # the statistics themselves: merged partials vs torch, per (sample, group)
    mr = ws[B * 64 * 32 * 3:].view(B, 32, 2)
    xg = buf.permute(0, 3, 1, 2).reshape(B, 32, -1).double()
    assert (mr[:, :, 0].double() - xg.mean(2)).abs().max() < 1e-4 * xg.abs().max()
    assert ((mr[:, :, 1].double() - (xg.var(2, unbiased=False) + 1e-5).rsqrt()) / mr[:, :, 1].double()).abs().max() < 2e-4


@pytest.mark.parametrize("prec,tol", [("bf16", 1.2e-2), ("fp16", 1.5e-3)])
@pytest.mark.parametrize("shape", [(3, 64, 64, 128, 128), (1, 256, 128, 64, 256), (2, 16, 16, 128, 384)])
def test_groupnorm_of_a_16bit_activation_with_epilogue_statistics(dev, prec, tol, shape):
    """The tensor between a ResBlock's two convolutions in the 16-bit modes (Engine.act_h): the conv writes it in the operand
    dtype ONLY, its GroupNorm partials come from the fp32 accumulators, and nlc_groupnorm(x_is_op=1) reads the 16-bit
    tensor.  Against F.group_norm (+ per-sample scale/shift, SiLU) of the fp32 conv output: the statistics are those of the
    unrounded tensor (1e-4), the activated output carries one extra operand rounding of its input."""
    from nlc_b200 import ops
    B, H, W, Cin, Cout = shape
    dt = _dt(prec)
    tdt = ops.OP_DTYPES[dt]
    g = torch.Generator().manual_seed(19)
    x = _rnd(torch.randn(B, Cin, H, W, generator=g).to(dev), dt)
    w1 = _rnd((torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5).to(dev), dt)
    b1 = (torch.randn(Cout, generator=g) + 3.0).to(dev)
    rv = torch.randn(B, Cout, generator=g).to(dev)
    ref_h = F.conv2d(x, w1, b1, padding=1) + rv[:, :, None, None]
    xa = ops.Act(x.permute(0, 2, 3, 1).contiguous().to(tdt))
    st = ops.GnStats(torch.zeros(B * H * W // 32, Cout // 4, 2, device=dev))
    h16 = ops.Act(torch.zeros(B, H, W, Cout, device=dev, dtype=tdt), 0, Cout, st)
    ops.conv_tc([xa], ops.taps3x3(0, 0, Cin), ops.pack_conv_weight(w1, dt), Cout, B, H, W, dt, bias=b1, rowvec=rv,
                out_op=h16, stats=True)
    assert _rel(h16.t.float().permute(0, 3, 1, 2), ref_h) < (5e-3 if prec == "bf16" else 6e-4)
    gam, bet = torch.randn(Cout, generator=g).to(dev), torch.randn(Cout, generator=g).to(dev)
    sc, sh = (0.3 * torch.randn(B, Cout, generator=g)).to(dev), torch.randn(B, Cout, generator=g).to(dev)
    ref = F.group_norm(ref_h, 32, gam, bet, eps=1e-5) * (1 + sc[:, :, None, None]) + sh[:, :, None, None]
    ref = F.silu(ref)
    y = ops.Act(torch.zeros(B, H, W, Cout, device=dev, dtype=tdt))
    ws = torch.zeros(ops.groupnorm_ws(B, H * W, Cout, 32), device=dev)
    ops.groupnorm(h16, 32, 1e-5, gam, bet, y, dt, ws, silu=True, scale=sc, shift=sh, use_stats=True)
    assert _rel(y.t.float().permute(0, 3, 1, 2), ref) < tol
    mr = ws[B * 64 * 32 * 3:].view(B, 32, 2)
    xg = ref_h.reshape(B, 32, -1).double()
    assert (mr[:, :, 0].double() - xg.mean(2)).abs().max() < 1e-4 * xg.abs().max()
    assert ((mr[:, :, 1].double() - (xg.var(2, unbiased=False) + 1e-5).rsqrt()) / mr[:, :, 1].double()).abs().max() < 2e-4


@pytest.mark.parametrize("prec,tol", [("bf16", 6e-3), ("tf32", 8e-4), ("fp16", 8e-4)])
@pytest.mark.parametrize("mode", [1, 2])
def test_groupnorm_with_fused_resample(dev, prec, tol, mode):
    """resample=1: nearest x2 of silu(gn(x)); resample=2: avg_pool2d(silu(gn(x)), 2) (ADM resblock_updown h_upd)."""
    from nlc_b200 import ops
    dt = _dt(prec)
    B, H, W, C = 3, 16, 8, 128
    g = torch.Generator().manual_seed(10)
    x = torch.randn(B, C, H, W, generator=g).to(dev) * 2 + 0.3
    gam, bet = torch.randn(C, generator=g).to(dev), torch.randn(C, generator=g).to(dev)
    sc, sh = torch.randn(B, C, generator=g).to(dev) * 0.3, torch.randn(B, C, generator=g).to(dev)
    a = F.silu(F.group_norm(x, 32, gam, bet, eps=1e-5) * (1 + sc[:, :, None, None]) + sh[:, :, None, None])
    ref = F.interpolate(a, scale_factor=2.0, mode="nearest") if mode == 1 else F.avg_pool2d(a, 2)
    Ho, Wo = ref.shape[2:]
    y = ops.Act(torch.zeros(B, Ho, Wo, C, device=dev, dtype=ops.OP_DTYPES[dt]))
    ws = torch.zeros(ops.groupnorm_ws(B, H * W, C, 32), device=dev)
    ops.groupnorm(ops.Act(x.permute(0, 2, 3, 1).contiguous()), 32, 1e-5, gam, bet, y, dt, ws, silu=True, scale=sc, shift=sh,
                  resample=mode)
    assert _rel(y.t.float().permute(0, 3, 1, 2), ref) < tol


@pytest.mark.parametrize("prec,tol", [("bf16", 6e-3), ("tf32", 6e-4), ("fp16", 6e-4)])
def test_input_conv_on_tensor_cores(dev, prec, tol):
    """im2col (per-sample input scale folded in) + K=64|32 GEMM == F.conv2d(x*scale, w, b); tolerance = one operand
    rounding of the 27-term dot product."""
    from nlc_b200 import ops
    dt = _dt(prec)
    g = torch.Generator().manual_seed(12)
    B, H, W, Cout = 3, 32, 16, 128
    x = torch.randn(B, 3, H, W, generator=g).to(dev) * 3
    sc = (torch.rand(B, generator=g) + 0.1).to(dev)
    w = (torch.randn(Cout, 3, 3, 3, generator=g) / 27 ** 0.5).to(dev)
    b = torch.randn(Cout, generator=g).to(dev)
    ref = F.conv2d(x * sc.view(-1, 1, 1, 1), w, b, padding=1)
    kp = 32 if prec == "tf32" else 64
    patches = ops.Act(torch.empty(B, H, W, kp, device=dev, dtype=ops.OP_DTYPES[dt]))
    ops.im2col_in(x, sc, patches, dt)
    out = ops.Act(torch.zeros(B, H, W, Cout, device=dev))
    ops.conv_tc([patches], [(0, 0, 0, 0, kp)], ops.pack_conv_in_weight(w, dt), Cout, B, H, W, dt, bias=b, out_f32=out)
    assert _rel(out.t.permute(0, 3, 1, 2), ref) < tol


PAIR_CASES = [
    # B, H, W, Cin, Cout, stride, pad, k     -- sizes at which nlc_conv_tc selects the CTA-pair (cta_group::2) kernel
    (40, 32, 32, 128, 256, 1, (1, 1, 1, 1), 3),    # 320 M tiles, BLOCK_N 256
    (149, 8, 16, 128, 256, 1, (1, 1, 1, 1), 3),    # odd tile count: the last pair has one CTA with nothing to store
    (80, 16, 16, 64, 128, 1, (1, 1, 1, 1), 3),     # BLOCK_N 128 pairs
    (48, 32, 32, 128, 512, 2, (0, 1, 0, 1), 3),    # stride 2, two N tiles
    (20, 64, 64, 192, 256, 1, (0, 0, 0, 0), 1),    # 1x1, K = 192 (three chunks)
    # W >= 128: an M tile is 128 pixels of one image row (the ADM 256x256 / 128x128 levels)
    (2, 128, 128, 128, 256, 1, (1, 1, 1, 1), 3),   # one tile per image row
    (1, 256, 256, 64, 128, 1, (1, 1, 1, 1), 3),    # two tiles per row, BLOCK_N 128
    (3, 128, 128, 192, 512, 1, (1, 1, 1, 1), 3),   # three K chunks per tap, two N tiles, odd batch
    (2, 128, 128, 128, 256, 1, (0, 0, 0, 0), 1),   # 1x1
]


@pytest.mark.parametrize("prec", ["bf16", "tf32", "fp16"])
@pytest.mark.parametrize("case", PAIR_CASES)
def test_conv_tc_cta_pair_kernel(dev, prec, case):
    """Same contract as test_conv_tc_matches_torch, plus GroupNorm partials, on the 256-row CTA-pair tiles."""
    from nlc_b200 import ops
    B, H, W, Cin, Cout, stride, pad, k = case
    dt = _dt(prec)
    g = torch.Generator().manual_seed(21)
    x = _rnd(torch.randn(B, Cin, H, W, generator=g).to(dev), dt)
    w = _rnd((torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).to(dev), dt)
    b = torch.randn(Cout, generator=g).to(dev)
    ref = F.conv2d(F.pad(x, pad), w, b, stride=stride)
    Ho, Wo = ref.shape[2:]
    rowvec = torch.randn(B, Cout, generator=g).to(dev)
    resid = torch.randn(B, Ho, Wo, Cout, generator=g).to(dev)
    ref = (ref + rowvec[:, :, None, None] + resid.permute(0, 3, 1, 2)) * 0.5
    tdt = ops.OP_DTYPES[dt]
    xa = ops.Act(x.permute(0, 2, 3, 1).contiguous().to(tdt))
    segs = [(0, kh - pad[2], kw - pad[0], 0, Cin) for kh in range(k) for kw in range(k)]
    st = ops.GnStats(torch.zeros(B * Ho * Wo // 32, Cout // 4, 2, device=dev)) if Ho * Wo >= 128 else None
    o32 = ops.Act(torch.full((B, Ho, Wo, Cout), float("nan"), device=dev), 0, Cout, st)
    oop = ops.Act(torch.zeros(B, Ho, Wo, Cout, device=dev, dtype=tdt))
    ops.conv_tc([xa], segs, ops.pack_conv_weight(w, dt), Cout, B, Ho, Wo, dt, stride=stride, bias=b, rowvec=rowvec,
                resid=ops.Act(resid), out_scale=0.5, out_f32=o32, out_op=oop, stats=st is not None)
    torch.cuda.synchronize()
    assert _rel(o32.t.permute(0, 3, 1, 2), ref) < 5e-5
    assert _rel(oop.t.float().permute(0, 3, 1, 2), ref) < (8e-3 if prec == "bf16" else 1e-3)
    if st is not None:
        blk = o32.t.view(B * Ho * Wo // 32, 32, Cout // 4, 4).permute(0, 2, 1, 3).reshape(-1, Cout // 4, 128).double()
        assert (st.t[:, :, 0].double() - blk.mean(2)).abs().max() < 1e-5 * blk.abs().max()
        m2 = ((blk - blk.mean(2, keepdim=True)) ** 2).sum(2)
        assert ((st.t[:, :, 1].double() - m2) / m2.clamp_min(1e-6)).abs().max() < 1e-3


def test_conv_tc_pair_kernel_fused_shortcut_wide_rows(dev):
    """The CTA-pair kernel with a mixed segment list at 128-pixel rows: 3x3 over one source + the ResNet block's 1x1
    shortcut over a channel slice of a second, wider buffer, bf16."""
    from nlc_b200 import ops
    from nlc_b200._lib import NLC_BF16
    B, H, C0, C1, Cout = 2, 128, 128, 192, 256
    g = torch.Generator().manual_seed(23)
    x0 = _rnd(torch.randn(B, C0, H, H, generator=g).to(dev), NLC_BF16)
    x1 = _rnd(torch.randn(B, C1, H, H, generator=g).to(dev), NLC_BF16)
    w0 = _rnd((torch.randn(Cout, C0, 3, 3, generator=g) / (C0 * 9) ** 0.5).to(dev), NLC_BF16)
    w1 = _rnd((torch.randn(Cout, C1, 1, 1, generator=g) / C1 ** 0.5).to(dev), NLC_BF16)
    ref = F.conv2d(x0, w0, padding=1) + F.conv2d(x1, w1)
    buf = torch.zeros(B, H, H, C1 + 64, device=dev, dtype=torch.bfloat16)
    buf[..., 64:] = x1.permute(0, 2, 3, 1).to(torch.bfloat16)
    out = ops.Act(torch.zeros(B, H, H, Cout, device=dev))
    ops.conv_tc([ops.Act(x0.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)), ops.Act(buf, 64, C1)],
                ops.taps3x3(0, 0, C0) + [(1, 0, 0, 0, C1)], ops.pack_conv_weight(w0, NLC_BF16, extra=w1), Cout, B, H, H,
                NLC_BF16, out_f32=out)
    torch.cuda.synchronize()
    assert _rel(out.t.permute(0, 3, 1, 2), ref) < 5e-5


@pytest.mark.parametrize("mode,shape", [(1, (3, 32, 32)), (2, (3, 32, 32)), (1, (40, 8, 8)), (2, (40, 8, 8)),
                                        (1, (2, 128, 128)), (2, (1, 128, 128))])
def test_conv_tc_resampled_residual(dev, mode, shape):
    """nlc_conv_desc.resid_mode: the residual is read at half (1: nearest x2) / double (2: 2x2 average) resolution by
    the epilogue, bit-identical to adding the separately resampled tensor (ADM resblock_updown x_upd)."""
    from nlc_b200 import _lib, ops
    from nlc_b200._lib import NLC_BF16
    B, H, W = shape
    C = 128
    # (both launches on the tap-per-tile kernel in one pass: the halo-slab kernel and the split-K path, which take the
    #  plain-residual one at these sizes, sum the same products in another order)
    _lib.check(_lib.lib().nlc_ctx_set(_lib.ctx(0), b"slab", 0))
    _lib.check(_lib.lib().nlc_ctx_set(_lib.ctx(0), b"splitk", 0))
    g = torch.Generator().manual_seed(31)
    x = _rnd(torch.randn(B, C, H, W, generator=g).to(dev), NLC_BF16)
    w = _rnd((torch.randn(C, C, 3, 3, generator=g) / (C * 9) ** 0.5).to(dev), NLC_BF16)
    rs = (H // 2, W // 2) if mode == 1 else (2 * H, 2 * W)
    resid = torch.randn(B, rs[0], rs[1], C, generator=g).to(dev)
    full = torch.empty(B, H, W, C, device=dev)
    ops.resample(ops.Act(resid), mode, ops.Act(full), None, NLC_BF16)
    xa = ops.Act(x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16))
    wp = ops.pack_conv_weight(w, NLC_BF16)
    a, b = ops.Act(torch.zeros(B, H, W, C, device=dev)), ops.Act(torch.zeros(B, H, W, C, device=dev))
    ops.conv_tc([xa], ops.taps3x3(0, 0, C), wp, C, B, H, W, NLC_BF16, resid=ops.Act(full), out_f32=a)
    ops.conv_tc([xa], ops.taps3x3(0, 0, C), wp, C, B, H, W, NLC_BF16, resid=ops.Act(resid), out_f32=b, resid_mode=mode)
    torch.cuda.synchronize()
    _lib.check(_lib.lib().nlc_ctx_set(_lib.ctx(0), b"slab", 1))
    _lib.check(_lib.lib().nlc_ctx_set(_lib.ctx(0), b"splitk", 1))
    assert torch.equal(a.t, b.t)
    ref = F.conv2d(x, w, padding=1) + full.permute(0, 3, 1, 2)
    assert _rel(b.t.permute(0, 3, 1, 2), ref) < 5e-5


@pytest.mark.parametrize("prec,tol", [("bf16", 6e-3), ("tf32", 8e-4)])
@pytest.mark.parametrize("shape", [(2, 16, 16, 128, 128), (3, 32, 32, 256, 256), (1, 16, 8, 64, 128)])
def test_upsample_conv_as_four_subpixel_phases(dev, prec, tol, shape):
    """"nearest x2, then 3x3 conv" (src/unet_ddim.py:58-74) computed at the low resolution: four phase convolutions with the
    summed 2x2 tap sets, each writing its quarter of the output in place (nlc_conv_desc.out_up) and the matching GroupNorm
    partials.  Reference: F.conv2d on the upsampled tensor with the phase weights' own rounding folded out (the phase sums
    are rounded to the operand dtype once, so the comparison uses them)."""
    from nlc_b200 import ops
    B, H, W, Cin, Cout = shape
    dt = _dt(prec)
    tdt = ops.OP_DTYPES[dt]
    g = torch.Generator().manual_seed(11)
    x = _rnd(torch.randn(B, Cin, H, W, generator=g).to(dev), dt)
    w = (torch.randn(Cout, Cin, 3, 3, generator=g) / (Cin * 9) ** 0.5).to(dev)
    b = (torch.randn(Cout, generator=g) + 3.0).to(dev)
    # reference with exactly the weights the kernel sees: per phase, the rounded 2x2 sums
    ref = torch.zeros(B, Cout, 2 * H, 2 * W, device=dev)
    packs = ops.upsample_phase_weights(w, dt)
    for (a, bb), pk in packs.items():
        k = pk.float().reshape(Cout, 2, 2, Cin).permute(0, 3, 1, 2)  # [Cout, Cin, 2, 2]
        offs_r = [d for d, _ in ops._UP_ROWS[a]]
        offs_c = [d for d, _ in ops._UP_ROWS[bb]]
        xp = F.pad(x, (1, 1, 1, 1))
        acc = b[None, :, None, None].expand(B, Cout, H, W).clone()
        for i, dh in enumerate(offs_r):
            for j, dw in enumerate(offs_c):
                acc = acc + torch.einsum("oc,nchw->nohw", k[:, :, i, j], xp[:, :, 1 + dh:1 + dh + H, 1 + dw:1 + dw + W])
        ref[:, :, a::2, bb::2] = acc
    # ... which is the upsample + 3x3 conv up to the rounding of the summed weights
    full = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, b, padding=1)
    assert _rel(ref, full) < (2e-2 if prec == "bf16" else 2e-3)
    xa = ops.Act(x.permute(0, 2, 3, 1).contiguous().to(tdt))
    buf = torch.full((B, 2 * H, 2 * W, Cout), float("nan"), device=dev)
    st = ops.GnStats(torch.zeros(B * 4 * H * W // 32, Cout // 4, 2, device=dev))
    o32 = ops.Act(buf, 0, Cout, st)
    oop = ops.Act(torch.zeros(B, 2 * H, 2 * W, Cout, device=dev, dtype=tdt))
    for (a, bb), pk in packs.items():
        ops.conv_tc([xa], ops.upsample_phase_taps(0, 0, Cin, a, bb), pk, Cout, B, H, W, dt, bias=b, out_f32=o32, out_op=oop,
                    stats=True, out_up=(a, bb))
    torch.cuda.synchronize()
    assert _rel(buf.permute(0, 3, 1, 2), ref) < 5e-5
    assert _rel(oop.t.float().permute(0, 3, 1, 2), ref) < (8e-3 if prec == "bf16" else 1e-3)
    gam, bet = torch.randn(Cout, generator=g).to(dev), torch.randn(Cout, generator=g).to(dev)
    want = F.silu(F.group_norm(buf.permute(0, 3, 1, 2), 32, gam, bet, eps=1e-5))
    y = ops.Act(torch.zeros(B, 2 * H, 2 * W, Cout, device=dev, dtype=tdt))
    ws = torch.zeros(ops.groupnorm_ws(B, 4 * H * W, Cout, 32), device=dev)
    ops.groupnorm(o32, 32, 1e-5, gam, bet, y, dt, ws, silu=True, use_stats=True)
    assert _rel(y.t.float().permute(0, 3, 1, 2), want) < tol


TMA_EPI_CASES = [
    # B, H, W, Cin, Cout, k, stride, residual, stats, channel-slice output
    (3, 64, 64, 128, 128, 3, 1, True, True, False),    # slab kernel (W = 64: box 32 x 1)
    (3, 16, 16, 128, 128, 3, 1, True, True, False),    # slab kernel, W = 16 (box 16 x 2), odd super-tile count
    (2, 32, 32, 128, 256, 3, 1, False, True, True),    # slab kernel, two n tiles, output = channel slice of a wider buffer
    (5, 4, 4, 256, 512, 3, 1, True, False, False),     # 8 images per M tile, ragged batch: rows past the batch are clipped
    (6, 8, 8, 256, 256, 3, 2, False, False, False),    # stride 2, two images per tile
    (3, 32, 32, 256, 768, 1, 1, False, False, False),  # qkv-shaped 1x1, N = 256 tiles
    (1, 256, 256, 64, 64, 3, 1, True, True, False),    # 128-pixel rows, N = 64
    (2, 2, 2, 512, 512, 3, 1, True, False, False),     # 32 images per tile
    (130, 1, 1, 256, 128, 1, 1, False, False, False),  # 1x1 spatial
    (9, 64, 64, 64, 128, 1, 1, True, True, False),     # CTA-pair kernel with an odd tile count
]


@pytest.mark.parametrize("prec", ["bf16", "fp16"])
@pytest.mark.parametrize("case", TMA_EPI_CASES)
def test_conv_tma_epilogue_equals_lsu_epilogue(dev, prec, case):
    """The 16-bit-only convolution epilogues - through TMA (ConvKParams.tma_epi 1: residual block by tensor load, output block
    by tensor store, SWIZZLE_64B staging) and with 256-bit global accesses from / to registers (2) - against the staged
    load/store-unit epilogue on the same launch: identical arithmetic in the same
    order, so outputs and GroupNorm partials must be BIT-equal; and nothing outside the output's channel slice / batch is
    touched (canary bands)."""
    from nlc_b200 import _lib, ops
    B, H, W, Cin, Cout, k, stride, with_resid, with_stats, sliced = case
    dt = _dt(prec)
    tdt = ops.OP_DTYPES[dt]
    g = torch.Generator().manual_seed(77)
    pad = k // 2
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    x = torch.randn(B, H, W, Cin, generator=g).to(dev).to(tdt)
    w = _rnd((torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).to(dev), dt)
    b = torch.randn(Cout, generator=g).to(dev)
    rowvec = torch.randn(B, Cout, generator=g).to(dev)
    resid = torch.randn(B, Ho, Wo, Cout, generator=g).to(dev).to(tdt) if with_resid else None
    segs = [(0, kh - pad, kw - pad, 0, Cin) for kh in range(k) for kw in range(k)]
    wp = ops.pack_conv_weight(w, dt)
    ctx = _lib.ctx(0)
    outs = {}
    try:
        for mode in (1, 2, 0):
            _lib.check(_lib.lib().nlc_ctx_set(ctx, b"tma_epi", mode))
            can_stats = with_stats and Ho * Wo >= 128
            st = ops.GnStats(torch.zeros(B * Ho * Wo // 32, (Cout + (64 if sliced else 0)) // 4, 2, device=dev)) if can_stats else None
            if sliced:  # the launch owns channels [32, 32 + Cout) of a buffer with 64 more; one extra image of canary behind
                buf = torch.full((B + 1, Ho, Wo, Cout + 64), 7.0, device=dev, dtype=tdt)
                oop = ops.Act(buf[:B], 32, Cout, st)
            else:
                buf = torch.full((B + 1, Ho, Wo, Cout), 7.0, device=dev, dtype=tdt)
                oop = ops.Act(buf[:B], 0, Cout, st)
            ops.conv_tc([ops.Act(x)], segs, wp, Cout, B, Ho, Wo, dt, stride=stride, bias=b, rowvec=rowvec,
                        resid=ops.Act(resid) if with_resid else None, out_scale=0.5, out_op=oop, stats=can_stats)
            torch.cuda.synchronize()
            outs[mode] = (buf.clone(), st.t.clone() if can_stats else None)
    finally:
        _lib.check(_lib.lib().nlc_ctx_set(ctx, b"tma_epi", 1))
    for mode in (1, 2):  # 1: TMA; 2: 256-bit global accesses straight from / to registers
        assert torch.equal(outs[mode][0].view(torch.int16), outs[0][0].view(torch.int16))
        if outs[mode][1] is not None:
            assert torch.equal(outs[mode][1], outs[0][1])
    buf = outs[1][0].float()
    assert torch.all(buf[B] == 7.0)
    if sliced:
        assert torch.all(buf[..., :32] == 7.0) and torch.all(buf[..., 32 + Cout:] == 7.0)
    xr = x.float().permute(0, 3, 1, 2)
    ref = F.conv2d(xr, w, b, stride=stride, padding=pad) + rowvec[:, :, None, None]
    if with_resid:
        ref = ref + resid.float().permute(0, 3, 1, 2)
    got = buf[:B, ..., 32:32 + Cout] if sliced else buf[:B]
    assert _rel(got.permute(0, 3, 1, 2), ref * 0.5) < (8e-3 if prec == "bf16" else 1e-3)


SPLITK_CASES = [
    # B, H, W, Cin, Cout, k, stride, fp32 residual?, 16-bit residual?, fp32 out?, operand out?
    (32, 4, 4, 512, 512, 3, 1, True, False, True, True),     # the c2 4x4 level: fp32 residual stream, both outputs
    (5, 4, 4, 256, 512, 3, 1, False, False, True, False),    # ragged batch
    (2, 2, 2, 512, 512, 3, 1, True, False, True, False),     # nine splits of eight chunks
    (8, 8, 8, 256, 256, 3, 2, False, False, False, True),    # stride 2, operand output only
    (4, 8, 8, 1024, 256, 1, 1, False, True, False, True),    # 1x1 with a 16-bit residual
    (130, 1, 1, 1024, 128, 1, 1, False, False, True, True),  # 1x1 spatial
]


@pytest.mark.parametrize("prec", ["fp16", "bf16", "tf32"])
@pytest.mark.parametrize("case", SPLITK_CASES)
def test_conv_split_k_equals_single_pass(dev, prec, case):
    """Deterministic split-K of the small-M launches (workspace + splitk_reduce_kernel): against the single-pass launch (same
    products, another fp32 summation order: 2e-5 of max|y|), against torch, and bit-reproducible over repeats."""
    from nlc_b200 import _lib, ops
    B, H, W, Cin, Cout, k, stride, r32, r16, o32_on, oop_on = case
    dt = _dt(prec)
    tdt = ops.OP_DTYPES[dt]
    if r16 and prec == "tf32":
        r16, r32 = False, True
    g = torch.Generator().manual_seed(31)
    pad = k // 2
    Ho, Wo = (H + 2 * pad - k) // stride + 1, (W + 2 * pad - k) // stride + 1
    x = _rnd(torch.randn(B, H, W, Cin, generator=g).to(dev), dt)
    w = _rnd((torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).to(dev), dt)
    b = torch.randn(Cout, generator=g).to(dev)
    rowvec = torch.randn(B, Cout, generator=g).to(dev)
    resid = torch.randn(B, Ho, Wo, Cout, generator=g).to(dev)
    resid = resid.to(tdt) if r16 else resid
    segs = [(0, kh - pad, kw - pad, 0, Cin) for kh in range(k) for kw in range(k)]
    wp = ops.pack_conv_weight(w, dt)
    ctx = _lib.ctx(0)
    outs = {}
    try:
        for mode in (1, 1, 0):
            _lib.check(_lib.lib().nlc_ctx_set(ctx, b"splitk", mode))
            o32 = ops.Act(torch.full((B, Ho, Wo, Cout), float("nan"), device=dev)) if o32_on else None
            oop = ops.Act(torch.zeros(B, Ho, Wo, Cout, device=dev, dtype=tdt)) if oop_on else None
            ops.conv_tc([ops.Act(x.to(tdt))], segs, wp, Cout, B, Ho, Wo, dt, stride=stride, bias=b, rowvec=rowvec,
                        resid=ops.Act(resid) if (r32 or r16) else None, out_scale=0.5, out_f32=o32, out_op=oop, relu=True)
            torch.cuda.synchronize()
            cur = (o32.t.clone() if o32_on else None, oop.t.float().clone() if oop_on else None)
            if mode == 1 and 1 in outs:  # second split-K run: bit-reproducible
                for a, c in zip(outs[1], cur):
                    assert a is None or torch.equal(a, c)
            outs[mode] = cur
    finally:
        _lib.check(_lib.lib().nlc_ctx_set(ctx, b"splitk", 1))
    ref = F.conv2d(x.permute(0, 3, 1, 2), w, b, stride=stride, padding=pad) + rowvec[:, :, None, None]
    if r32 or r16:
        ref = ref + resid.float().permute(0, 3, 1, 2)
    ref = torch.relu(ref * 0.5)
    if o32_on:
        assert _rel(outs[1][0], outs[0][0]) < 2e-5
        assert _rel(outs[1][0].permute(0, 3, 1, 2), ref) < 5e-5
    if oop_on:
        tol = 8e-3 if prec == "bf16" else 1e-3
        assert _rel(outs[1][1], outs[0][1]) < tol and _rel(outs[1][1].permute(0, 3, 1, 2), ref) < tol
