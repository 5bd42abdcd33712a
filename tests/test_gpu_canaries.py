"""Out-of-bounds guards (compute-sanitizer is not available on the GPU pool): every kernel writes its output into the
middle of a larger buffer filled with a sentinel, and the guard bands on both sides must come back untouched.  Covers
the kernels added late in round 1 (small-image GroupNorm, fused attention at both head dimensions, the tensor-core
output convolution and its NCHW head copy, linear, image metrics, the resampled-residual conv epilogue)."""
import pytest
import torch

pytestmark = pytest.mark.gpu
dev = torch.device("cuda:0")
SENT = 12345.0
GUARD = 4096


def _sent(dtype):
    return 0xA5 if dtype == torch.uint8 else SENT


def guarded(shape, dtype=torch.float32):
    n = 1
    for s in shape:
        n *= s
    buf = torch.full((n + 2 * GUARD,), _sent(dtype), device=dev, dtype=dtype)
    return buf, buf[GUARD:GUARD + n].view(*shape)


def intact(buf, n):
    s = torch.tensor(_sent(buf.dtype), dtype=buf.dtype)
    return bool((buf[:GUARD] == s).all() and (buf[GUARD + n:] == s).all())


@pytest.mark.parametrize("prec", ["bf16", "fp16", "tf32"])
def test_groupnorm_small_and_large(prec):
    from nlc_b200 import ops
    from nlc_b200.engine import PRECISIONS
    dt = PRECISIONS[prec]
    for (B, H, W, C) in [(3, 8, 8, 256), (5, 4, 4, 512), (2, 2, 2, 1024), (2, 32, 32, 128)]:
        x = torch.randn(B, H, W, C, device=dev)
        buf, y = guarded((B, H, W, C), ops.OP_DTYPES[dt])
        ws = torch.zeros(ops.groupnorm_ws(B, H * W, C, 32), device=dev)
        ops.groupnorm(ops.Act(x), 32, 1e-5, torch.ones(C, device=dev), torch.zeros(C, device=dev), ops.Act(y), dt, ws)
        torch.cuda.synchronize()
        assert intact(buf, y.numel()) and torch.isfinite(y.float()).all() and (y.float() != SENT).all()


@pytest.mark.parametrize("case", [(3, 64, 4, 64), (2, 256, 16, 64), (1, 1024, 8, 64), (5, 256, 1, 256), (3, 64, 2, 256)])
def test_fused_attention(case):
    from nlc_b200 import ops
    from nlc_b200._lib import NLC_BF16
    B, T, heads, dh = case
    C = heads * dh
    side = 1 << ((T.bit_length() - 1) // 2)
    qkv = torch.randn(B, side, T // side, 3 * C, device=dev).to(torch.bfloat16)
    buf, out = guarded((B, side, T // side, C), torch.bfloat16)
    nws = max(ops.attention_ws(NLC_BF16, B, T, heads, dh), 16)
    wbuf, ws = guarded((nws,), torch.uint8)
    ops.attention(ops.Act(qkv), NLC_BF16, 0, C, 2 * C, dh, heads, dh, dh ** -0.5, ops.Act(out), ws)
    torch.cuda.synchronize()
    assert intact(buf, out.numel()) and torch.isfinite(out.float()).all()
    assert intact(wbuf, nws)


def test_conv_out_head_linear_metrics_and_resampled_residual():
    from nlc_b200 import ops
    from nlc_b200._lib import NLC_BF16, lib, ctx
    import ctypes as C
    B, H, W, Cin = 3, 32, 32, 128
    # tensor-core conv_out: padded 64-channel tile -> NCHW head copy
    a = ops.Act(torch.randn(B, H, W, Cin, device=dev).to(torch.bfloat16))
    wp, bp = ops.pack_conv_out_weight(torch.randn(6, Cin, 3, 3, device=dev) * 0.05, torch.randn(6, device=dev), NLC_BF16)
    tbuf, tmp = guarded((B, H, W, 64))
    ops.conv_tc([a], ops.taps3x3(0, 0, Cin), wp, 64, B, H, W, NLC_BF16, bias=bp, out_f32=ops.Act(tmp))
    obuf, out = guarded((B, 6, H, W))
    ops.nhwc_head_to_nchw(ops.Act(tmp), 6, out)
    torch.cuda.synchronize()
    assert intact(tbuf, tmp.numel()) and intact(obuf, out.numel()) and (out != SENT).all()
    assert torch.equal(out, tmp[..., :6].permute(0, 3, 1, 2))
    # linear, ragged sizes
    x = torch.randn(37, 516, device=dev)
    Wm = torch.randn(1001, 516, device=dev)
    ybuf, y = guarded((37, 1001))
    ops.linear(x, Wm, None, y)
    torch.cuda.synchronize()
    assert intact(ybuf, y.numel()) and (y != SENT).all()
    # image metrics
    n = 3 * 17 * 9
    mbuf, m = guarded((5,))
    lbuf, l1 = guarded((5,))
    sbuf, s01 = guarded((5, n))
    xs, xo = torch.randn(5, n, device=dev), torch.rand(5, n, device=dev)
    rc = lib().nlc_image_metrics(ctx(0), xs.data_ptr(), xo.data_ptr(), 5, n, C.c_void_p(s01.data_ptr()), m.data_ptr(),
                                 l1.data_ptr(), C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert rc == 0 and intact(mbuf, 5) and intact(lbuf, 5) and intact(sbuf, s01.numel())
    # conv with the residual read at half / double resolution
    for mode, rs in ((1, (H // 2, W // 2)), (2, (2 * H, 2 * W))):
        resid = torch.randn(B, rs[0], rs[1], Cin, device=dev)
        w3 = ops.pack_conv_weight(torch.randn(Cin, Cin, 3, 3, device=dev) * 0.03, NLC_BF16)
        fbuf, f = guarded((B, H, W, Cin))
        ops.conv_tc([a], ops.taps3x3(0, 0, Cin), w3, Cin, B, H, W, NLC_BF16, resid=ops.Act(resid), out_f32=ops.Act(f),
                    resid_mode=mode)
        torch.cuda.synchronize()
        assert intact(fbuf, f.numel()) and (f != SENT).all()


@pytest.mark.parametrize("name", ["inpainting", "colorization", "sr2", "sr4", "whcs", "deblur", "denoising", "cs"])
def test_operator_and_ddnm_kernels_stay_inside_their_buffers(name):
    """The 128-bit operator kernels, the warp-shuffle FWHT with its fused epilogues, block CS, the fused DDNM / DDNM+ step,
    SSIM: outputs and the workspace sit between guard bands, called through the C ABI with odd batch sizes."""
    import ctypes as C
    from nlc_b200 import _lib, svd_operators as P
    from oracle import operators as O
    R, Ch, B = 64, 3, 3
    gen = torch.Generator().manual_seed(5)
    mask = (torch.rand(R, R, generator=gen) > 0.3).float()
    mr = torch.nonzero(mask.reshape(-1) == 0).long().reshape(-1) * 3
    missing = torch.cat([mr, mr + 1, mr + 2])
    op = {"inpainting": lambda: P.Inpainting(Ch, R, missing, dev), "colorization": lambda: P.Colorization(R, dev),
          "sr2": lambda: P.SuperResolution(Ch, R, 2, dev), "sr4": lambda: P.SuperResolution(Ch, R, 4, dev),
          "whcs": lambda: P.WalshHadamardCS(Ch, R, 4, torch.randperm(R * R, generator=gen), dev),
          "deblur": lambda: P.Deblurring(O.gauss_kernel(), Ch, R, dev), "denoising": lambda: P.Denoising(Ch, R, dev),
          "cs": lambda: P.CS(Ch, R, 0.25, dev, V_small=O.hadamard_basis(1024, 3))}[name]()
    L, st = _lib.lib(), C.c_void_p(torch.cuda.current_stream().cuda_stream)
    d = Ch * R * R
    nws = max(int(L.nlc_op_ws(op._h, B)), 16)
    wbuf, ws = guarded((nws,), torch.uint8)
    wsp = C.c_void_p(ws.data_ptr())
    x = torch.randn(B, d, device=dev)
    ybuf, y = guarded((B, op.ydim))
    _lib.check(L.nlc_op_A(op._h, x.data_ptr(), B, y.data_ptr(), wsp, st))
    outs = []
    for fn, src in ((L.nlc_op_At, y), (L.nlc_op_Apinv, y)):
        buf, o = guarded((B, d))
        _lib.check(fn(op._h, src.data_ptr(), B, o.data_ptr(), wsp, st))
        outs.append((buf, o))
    buf, o = guarded((B, d))
    _lib.check(L.nlc_op_Apinv_eta(op._h, y.data_ptr(), B, 0.1, o.data_ptr(), wsp, st))
    outs.append((buf, o))
    buf, o = guarded((B, d))
    _lib.check(L.nlc_op_project(op._h, x.data_ptr(), y.data_ptr(), B, o.data_ptr(), wsp, st))
    outs.append((buf, o))
    et, z = torch.randn(B, 2 * d, device=dev), torch.randn(B, d, device=dev)
    plus_modes = (0,) if name == "cs" else (0, 1)
    for plus in plus_modes:
        b0, x0 = guarded((B, d))
        b1, xn = guarded((B, d))
        _lib.check(L.nlc_ddnm_step(op._h, x.data_ptr(), et.data_ptr(), 2 * d, z.data_ptr(), y.data_ptr(), B, 0.5, 0.6,
                                   0.85, 0.1, plus, x0.data_ptr(), xn.data_ptr(), wsp, st))
        outs += [(b0, x0), (b1, xn)]
    if name != "cs":
        coef = _lib.DdnmCoef(a=0.8, sigma_t=0.3, sigma_y=0.2, eta=0.85)
        for two in (False, True):
            buf, o = guarded((B, d))
            if two:
                _lib.check(L.nlc_op_lambda_noise(op._h, x.data_ptr(), z.data_ptr(), B, C.byref(coef), o.data_ptr(), wsp, st))
            else:
                _lib.check(L.nlc_op_lambda(op._h, x.data_ptr(), B, C.byref(coef), o.data_ptr(), wsp, st))
            outs.append((buf, o))
    torch.cuda.synchronize()
    assert intact(ybuf, y.numel()) and intact(wbuf, nws) and (y != SENT).all()
    for buf, o in outs:
        assert intact(buf, o.numel()) and torch.isfinite(o).all() and (o != SENT).all()


def test_ssim_stays_inside_its_buffers():
    import ctypes as C
    from nlc_b200 import _lib
    L, st = _lib.lib(), C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for (B, H, W) in ((3, 40, 52), (2, 16, 32), (1, 7, 5)):
        a, b = torch.rand(B, 3, H, W, device=dev), torch.rand(B, 3, H, W, device=dev)
        nws = int(L.nlc_ssim3d_ws(B, H, W))
        wbuf, ws = guarded((nws,), torch.uint8)
        obuf, out = guarded((B,))
        _lib.check(L.nlc_ssim3d(_lib.ctx(0), a.data_ptr(), b.data_ptr(), B, H, W, C.c_void_p(ws.data_ptr()), out.data_ptr(), st))
        torch.cuda.synchronize()
        assert intact(wbuf, nws) and intact(obuf, B) and ((out > -1) & (out <= 1)).all()


def test_training_kernels_stay_inside_their_buffers():
    import ctypes as C
    from nlc_b200 import _lib
    L, st = _lib.lib(), C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for (B, d) in ((3, 3 * 16 * 16), (2, 1001)):
        x0, noise, extra = (torch.randn(B, d, device=dev) for _ in range(3))
        eta1, eta2, ab = torch.rand(B, device=dev), torch.rand(B, device=dev), torch.rand(B, device=dev) * 0.9 + 0.05
        b1, noisy = guarded((B, d))
        b2, nn = guarded((B, d))
        b3, dist = guarded((B,))
        _lib.check(L.nlc_train_prepare(_lib.ctx(0), x0.data_ptr(), noise.data_ptr(), extra.data_ptr(), eta1.data_ptr(),
                                       eta2.data_ptr(), ab.data_ptr(), B % 2, B, d, noisy.data_ptr(),
                                       C.c_void_p(nn.data_ptr()), dist.data_ptr(), st))
        torch.cuda.synchronize()
        assert intact(b1, B * d) and intact(b2, B * d) and intact(b3, B) and torch.isfinite(noisy).all() and (dist > 0).all()
    for n in (4096, 4099):
        bufs = [guarded((n,)) for _ in range(4)]
        for _, v in bufs:
            v.copy_(torch.randn(n, device=dev).abs())
        g = torch.randn(n, device=dev)
        _lib.check(L.nlc_adamw_ema_step(_lib.ctx(0), bufs[0][1].data_ptr(), g.data_ptr(), bufs[1][1].data_ptr(),
                                        bufs[2][1].data_ptr(), bufs[3][1].data_ptr(), n, 1e-3, 0.9, 0.999, 1e-8, 0.01, 1, 0.99,
                                        1.0, st))
        torch.cuda.synchronize()
        for buf, v in bufs:
            assert intact(buf, n) and torch.isfinite(v).all()
