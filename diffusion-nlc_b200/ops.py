"""Python-side wrappers of the C ABI: thin argument marshalling from torch tensors (device memory +
current stream) to libnlc_b200 entry points.  No arithmetic happens here."""
import ctypes as C

import torch

from . import _lib
from ._lib import NLC_BF16, NLC_F16, NLC_F32, NLC_F32X3

OP_DTYPES = {NLC_BF16: torch.bfloat16, NLC_F16: torch.float16, NLC_F32: torch.float32, NLC_F32X3: torch.float32}


class _Stats:
    """Launch accounting for bench.py: `launches` counts kernels enqueued through this module; when
    `conv_timer` is a list, every nlc_conv_tc launch is bracketed by CUDA events on the launching stream and
    (flops, start, stop) is appended."""
    launches = 0
    graph_replays = 0  # CUDA-graph replays of a captured timestep (each adds the captured launch count to `launches`)
    conv_timer = None
    op_timer = None  # list -> every wrapper below appends (name, start_event, stop_event, flops)


STATS = _Stats()


def _timed(name, name_detail=None):
    """Profiling aid (scripts/step_profile.py): bracket the wrapped launch with CUDA events on the launching
    stream when STATS.op_timer is a list.  No effect otherwise."""
    def deco(fn):
        def wrapped(*a, **k):
            t = STATS.op_timer
            if t is None:
                return fn(*a, **k)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **k)
            e1.record()
            t.append((name if not callable(name_detail) else name + name_detail(*a, **k), e0, e1, 0.0))
            return r
        wrapped.__name__, wrapped.__doc__ = fn.__name__, fn.__doc__
        return wrapped
    return deco


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ctx(t):
    return _lib.ctx(t.device.index if t.device.index is not None else torch.cuda.current_device())


def _p(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class Act:
    """NHWC activation view: channels [c0, c0+C) of a contiguous [B,H,W,Ctot] tensor.

    `ld` (= Ctot) is the pixel pitch in elements, so a view can be one half of a concatenation buffer
    (torch.cat at src/unet_ddim.py:353 never materialises separately)."""

    __slots__ = ("t", "c0", "C", "stats")

    def __init__(self, t, c0=0, C=None, stats=None):
        assert t.dim() == 4 and t.is_contiguous(), "Act wants a contiguous [B,H,W,C] tensor"
        self.t, self.c0 = t, c0
        self.C = t.shape[3] - c0 if C is None else C
        self.stats = stats  # GnStats of the underlying buffer (fp32 activations that a GroupNorm may read)

    B = property(lambda s: s.t.shape[0])
    H = property(lambda s: s.t.shape[1])
    W = property(lambda s: s.t.shape[2])
    ld = property(lambda s: s.t.shape[3])
    dtype = property(lambda s: s.t.dtype)
    ptr = property(lambda s: s.t.data_ptr() + s.c0 * s.t.element_size())

    def slice(self, c0, C):
        return Act(self.t, self.c0 + c0, C, self.stats)

    def dense(self):
        """[B,H,W,C] torch view (for tests)."""
        return self.t[..., self.c0:self.c0 + self.C]


class GnStats:
    """GroupNorm partial statistics of one fp32 NHWC buffer, written by the epilogues of the convolutions that
    produce it (nlc_conv_desc.stats): `t` is [B*H*W/32, Ctot/4, 2] (mean, M2 per 32 pixels x 4 channels);
    `covered` records, at plan time, which channel ranges have a producer that writes them."""

    __slots__ = ("t", "covered")

    def __init__(self, t):
        self.t, self.covered = t, []

    def covers(self, c0, C):
        need = c0
        for lo, hi in sorted(self.covered):
            if lo <= need < hi:
                need = hi
        return need >= c0 + C

    @staticmethod
    def eligible(B, H, W, C):
        """What the conv epilogue needs: a tile inside one image (H*W >= 128) and whole 4-channel blocks."""
        return H * W >= 128 and (H * W) % 32 == 0 and C % 4 == 0


def round_tf32_(w):
    """In-place fp32 -> tf32 (round to nearest, ties away): what cvt.rna.tf32.f32 does on the device."""
    bits = w.view(torch.int32)
    bits.add_(0x1000).bitwise_and_(~0x1FFF)
    return w


def pack_conv_weight(w, op_dtype, extra=None):
    """torch [Cout,Cin,kh,kw] -> K-major [Cout, kh*kw*Cin] (tap-major, channel-minor), optionally followed
    by more K columns (a fused 1x1 shortcut, [Cout, Cx])."""
    Cout = w.shape[0]
    k = w.detach().float().permute(0, 2, 3, 1).reshape(Cout, -1)
    if extra is not None:
        k = torch.cat([k, extra.detach().float().reshape(Cout, -1)], dim=1)
    k = k.contiguous()
    if op_dtype in (NLC_BF16, NLC_F16):
        return k.to(OP_DTYPES[op_dtype]).contiguous()
    if op_dtype == NLC_F32X3:
        return k.clone()  # split into tf32 hi + lo by the kernel
    return round_tf32_(k.clone())


# "nearest-neighbour x2 upsample, then 3x3 conv (padding 1)" at the low resolution: output pixel (2i+a, 2j+b) only sees the
# low-resolution rows {i-1, i} (a = 0) or {i, i+1} (a = 1), with the kernel rows that land on the same source row summed
_UP_ROWS = {0: ((-1, (0,)), (0, (1, 2))), 1: ((0, (0, 1)), (1, (2,)))}  # phase -> ((source offset, kernel rows), ...)


def upsample_phase_weights(w, op_dtype):
    """torch [Cout,Cin,3,3] -> {(a, b): packed [Cout, 4*Cin]} for the four sub-pixel phases (taps in `upsample_phase_taps`
    order).  The sums are taken in fp32 before the operand rounding."""
    w = w.detach().float()
    out = {}
    for a in (0, 1):
        for b in (0, 1):
            k = torch.zeros(w.shape[0], w.shape[1], 2, 2, dtype=torch.float32, device=w.device)
            for i, (_, rows) in enumerate(_UP_ROWS[a]):
                for j, (_, cols) in enumerate(_UP_ROWS[b]):
                    for kh in rows:
                        for kw in cols:
                            k[:, :, i, j] += w[:, :, kh, kw]
            out[(a, b)] = pack_conv_weight(k, op_dtype)
    return out


def upsample_phase_taps(src, c0, nch, a, b):
    """The four K segments (source offsets) of phase (a, b), in the order `upsample_phase_weights` packs them."""
    return [(src, dh, dw, c0, nch) for dh, _ in _UP_ROWS[a] for dw, _ in _UP_ROWS[b]]


def taps3x3(src, c0, nch, pad=1):
    """Nine K segments of a 3x3 convolution over channels [c0,c0+nch) of source `src`."""
    return [(src, kh - pad, kw - pad, c0, nch) for kh in range(3) for kw in range(3)]


def conv_tc(srcs, segs, weight, Cout, B, Ho, Wo, op_dtype, stride=1, bias=None, rowvec=None, resid=None,
            out_scale=1.0, out_f32=None, out_op=None, stats=False, resid_mode=0, out_up=None, relu=False):
    """Tensor-core implicit GEMM (nlc_conv_tc). srcs: list[Act]; segs: list of (src, dh, dw, c0, nch);
    resid/out_f32/out_op: Act or None; rowvec: [B, Cout] fp32 tensor."""
    d = _lib.ConvDesc()
    d.dtype = op_dtype
    d.nsrc = len(srcs)
    for i, a in enumerate(srcs):
        assert a.dtype == OP_DTYPES[op_dtype], "operand dtype mismatch"
        d.src[i] = _lib.Operand(a.ptr, a.B, a.H, a.W, a.C, a.ld)
    d.nseg = len(segs)
    for i, s in enumerate(segs):
        d.seg[i] = _lib.KSeg(*s)
    d.weight = weight.data_ptr()
    d.Cout, d.stride, d.B, d.Ho, d.Wo = Cout, stride, B, Ho, Wo
    d.bias = bias.data_ptr() if bias is not None else None
    if rowvec is not None:
        d.rowvec, d.ld_rowvec = rowvec.data_ptr(), rowvec.stride(0)
    if resid is not None:
        d.resid, d.ld_resid = resid.ptr, resid.ld
        d.resid_is_op = 0 if resid.dtype == torch.float32 else 1  # (16-bit residual stream: the operand dtype)
        assert not d.resid_is_op or resid.dtype == OP_DTYPES[op_dtype], "a 16-bit residual must be in the operand dtype"
        d.resid_mode = resid_mode  # 1 / 2: the residual is at half / double resolution (nearest x2 / 2x2 average)
    d.out_scale = out_scale
    d.act = 1 if relu else 0
    if out_up is not None:  # (a, b): this launch is one phase of "nearest x2, then 3x3" computed at the low resolution
        d.out_up = 1 + 2 * out_up[0] + out_up[1]
    if out_f32 is not None:
        d.out_f32, d.ld_out_f32 = out_f32.ptr, out_f32.ld
    if out_op is not None:
        d.out_op, d.ld_out_op = out_op.ptr, out_op.ld
    if stats:  # GroupNorm partials of the output (taken from the fp32 accumulators), at its channel offset inside the
        # buffer's stats tensor; the holder hangs on the fp32 output, or on the operand copy when that is the only one
        holder = out_f32 if (out_f32 is not None and out_f32.stats is not None) else out_op
        st = holder.stats.t
        d.stats, d.stats_nblk = st.data_ptr() + (holder.c0 // 4) * 8, st.shape[1]
    timer, optimer = STATS.conv_timer, STATS.op_timer
    if timer is not None or optimer is not None:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
    _lib.check(_lib.lib().nlc_conv_tc(_ctx(srcs[0].t), C.byref(d), _stream()))
    STATS.launches += 1
    if timer is not None or optimer is not None:
        e1.record()
        flops = 2.0 * B * Ho * Wo * Cout * sum(sg[4] for sg in segs)
        if out_up is not None:
            # a sub-pixel phase executes 4 of the 9 taps' multiplies of the "upsample, then 3x3" it computes: the bench counts
            # the ALGORITHMIC work of the layer (2 * pixels * Cout * 9 Cin), as for every other launch
            flops *= 9.0 / 4.0
        if timer is not None:
            timer.append((flops, e0, e1))
        if optimer is not None:
            optimer.append(("conv_tc %dx%d K%d N%d s%d%s" % (Ho, Wo, sum(sg[4] for sg in segs), Cout, stride,
                                                           " +f32" if out_f32 is not None else ""), e0, e1, flops))


@_timed("conv_in_nchw")
def conv_in_nchw(x_nchw, in_scale, weight, bias, out_f32, out_op, op_dtype):
    """Direct 3x3 conv on the sampler's NCHW fp32 image -> NHWC (nlc_conv_in_nchw)."""
    B, Cin, H, W = x_nchw.shape
    Cout = weight.shape[0]
    _lib.check(_lib.lib().nlc_conv_in_nchw(
        _ctx(x_nchw), _p(x_nchw), _p(in_scale), B, Cin, H, W, _p(weight), _p(bias), Cout,
        C.c_void_p(out_f32.ptr) if out_f32 is not None else None, out_f32.ld if out_f32 is not None else 0,
        C.c_void_p(out_op.ptr) if out_op is not None else None, out_op.ld if out_op is not None else 0,
        op_dtype, _stream()))
    STATS.launches += 1


@_timed("im2col_in")
def im2col_in(x_nchw, in_scale, patches, op_dtype):
    """NCHW fp32 image -> [B,H,W,64|32] K-major 3x3 patches in the operand dtype (nlc_im2col_in)."""
    B, Cin, H, W = x_nchw.shape
    _lib.check(_lib.lib().nlc_im2col_in(_ctx(x_nchw), _p(x_nchw), _p(in_scale), B, Cin, H, W, C.c_void_p(patches.ptr),
                                        op_dtype, _stream()))
    STATS.launches += 1


def pack_conv_out_weight(w, bias, op_dtype, n_pad=64):
    """torch [Cout<=8, Cin, 3, 3] (+bias) -> ([n_pad, 9*Cin] K-major operand weights, [n_pad] fp32 bias), zero rows
    past Cout: the output convolution as one 64-channel tensor-core tile."""
    Cout, Cin = w.shape[0], w.shape[1]
    wp = torch.zeros(n_pad, Cin, 3, 3, dtype=torch.float32, device=w.device)
    wp[:Cout] = w.detach().float()
    bp = torch.zeros(n_pad, dtype=torch.float32, device=w.device)
    if bias is not None:
        bp[:Cout] = bias.detach().float()
    return pack_conv_weight(wp, op_dtype), bp


def pack_conv_in_weight(w, op_dtype):
    """torch [Cout,Cin,3,3] -> [Cout, 64|32]: column tap*Cin + ci, zero-padded to one 128-byte K row."""
    Cout, Cin = w.shape[0], w.shape[1]
    kp = 64 if op_dtype in (NLC_BF16, NLC_F16) else 32
    k = torch.zeros(Cout, kp, dtype=torch.float32, device=w.device)
    k[:, :9 * Cin] = w.detach().float().permute(0, 2, 3, 1).reshape(Cout, -1)
    if op_dtype in (NLC_BF16, NLC_F16):
        return k.to(OP_DTYPES[op_dtype]).contiguous()
    return k if op_dtype == NLC_F32X3 else round_tf32_(k)


@_timed("conv_out_nchw")
def conv_out_nchw(x_op, op_dtype, weight, bias, out_nchw):
    """Direct 3x3 conv NHWC operand -> NCHW fp32 (nlc_conv_out_nchw)."""
    Cout = weight.shape[0]
    _lib.check(_lib.lib().nlc_conv_out_nchw(
        _ctx(x_op.t), C.c_void_p(x_op.ptr), op_dtype, x_op.ld, x_op.B, x_op.C, x_op.H, x_op.W, _p(weight), _p(bias),
        Cout, _p(out_nchw), _stream()))
    STATS.launches += 1


@_timed("nhwc_head_to_nchw")
def nhwc_head_to_nchw(x, n_ch, out_nchw):
    """First n_ch (<= 8) channels of an fp32 NHWC Act -> NCHW fp32 (nlc_nhwc_head_to_nchw)."""
    _lib.check(_lib.lib().nlc_nhwc_head_to_nchw(_ctx(x.t), C.c_void_p(x.ptr), x.ld, x.B, x.H, x.W, n_ch, _p(out_nchw),
                                                _stream()))
    STATS.launches += 1


def groupnorm_ws(B, HW, C, groups):
    return int(_lib.lib().nlc_groupnorm_ws(B, HW, C, groups))


@_timed("groupnorm", lambda x, *a, **k: " %dx%d C%d%s%s" % (x.H, x.W, x.C, " fused" if k.get("use_stats") else "",
                                                         " rs%d" % k["resample"] if k.get("resample") else ""))
def groupnorm(x, groups, eps, gamma, beta, y_op, op_dtype, ws, silu=True, scale=None, shift=None, use_stats=False,
              resample=0):
    """GroupNorm (+scale/shift, +SiLU) of NHWC `x` (Act: fp32, or the 16-bit operand dtype with use_stats) into operand
    `y_op` (Act).  use_stats: merge the partials the producing convolutions wrote (x.stats) instead of a statistics pass;
    resample 1/2: write the activated tensor nearest-x2 upsampled / 2x2 average pooled."""
    x_is_op = 0 if x.dtype == torch.float32 else 1
    assert not x_is_op or (use_stats and x.dtype == OP_DTYPES[op_dtype]), "16-bit GroupNorm input needs fused statistics"
    ld_ss = scale.stride(0) if scale is not None else 0
    st_ptr, st_nblk = None, 0
    if use_stats:
        st = x.stats.t
        st_ptr, st_nblk = C.c_void_p(st.data_ptr() + (x.c0 // 4) * 8), st.shape[1]
    _lib.check(_lib.lib().nlc_groupnorm(
        _ctx(x.t), C.c_void_p(x.ptr), x_is_op, x.ld, x.B, x.H, x.W, x.C, groups, eps, _p(gamma), _p(beta), _p(scale),
        _p(shift), ld_ss, 1 if silu else 0, st_ptr, st_nblk, resample, C.c_void_p(y_op.ptr), y_op.ld, op_dtype, _p(ws),
        _stream()))
    small = not use_stats and not resample and x.H * x.W * ((x.C // groups) // 4) <= 1024  # one fused launch
    STATS.launches += 1 if small else (2 if use_stats else 3)


@_timed("resample")
def resample(x, mode, y_f32, y_op, op_dtype):
    """mode 0 copy/cast, 1 nearest x2, 2 avgpool 2x2; x fp32 Act -> fp32 and/or operand Act."""
    _lib.check(_lib.lib().nlc_resample(
        _ctx(x.t), C.c_void_p(x.ptr), x.ld, x.B, x.H, x.W, x.C, mode,
        C.c_void_p(y_f32.ptr) if y_f32 is not None else None, y_f32.ld if y_f32 is not None else 0,
        C.c_void_p(y_op.ptr) if y_op is not None else None, y_op.ld if y_op is not None else 0, op_dtype, _stream()))
    STATS.launches += 1


@_timed("resample_op")
def resample_op(x_op, mode, y_op, op_dtype):
    """mode 1 nearest x2, 2 avgpool 2x2 on an operand-dtype Act."""
    _lib.check(_lib.lib().nlc_resample_op(_ctx(x_op.t), C.c_void_p(x_op.ptr), op_dtype, x_op.ld, x_op.B, x_op.H, x_op.W,
                                          x_op.C, mode, C.c_void_p(y_op.ptr), y_op.ld, _stream()))
    STATS.launches += 1


def attention_ws(op_dtype, B, T, heads, dh):
    return int(_lib.lib().nlc_attention_ws(op_dtype, B, T, heads, dh))


@_timed("attention", lambda qkv, op_dtype, q_off, k_off, v_off, head_stride, heads, dh, *a, **k: " T%d heads%d dh%d" % (
    qkv.H * qkv.W, heads, dh))
def attention(qkv, op_dtype, q_off, k_off, v_off, head_stride, heads, dh, scale, out, ws):
    """qkv: Act [B,H,W,ld]; out: Act [B,H,W,heads*dh] (operand dtype)."""
    T = qkv.H * qkv.W
    _lib.check(_lib.lib().nlc_attention(
        _ctx(qkv.t), C.c_void_p(qkv.ptr), op_dtype, qkv.ld, q_off, k_off, v_off, head_stride, qkv.B, T, heads, dh,
        scale, C.c_void_p(out.ptr), out.ld, _p(ws), _stream()))
    fused = op_dtype in (NLC_BF16, NLC_F16) and dh in (64, 256) and T % 64 == 0 and 64 <= T <= 1024
    STATS.launches += 2 if fused else (4 if T >= 128 else 1)


@_timed("linear")
def linear(x, W, bias, y, act_in=0, act_out=0):
    """y = act_out(act_in(x) @ W.T + bias); x [B,K], W [N,K], y [B,N] fp32 (rows may be strided)."""
    B, K = x.shape
    N = W.shape[0]
    assert W.is_contiguous() and W.shape[1] == K and x.stride(1) == 1 and y.stride(1) == 1
    _lib.check(_lib.lib().nlc_linear(_ctx(x), _p(x), x.stride(0), B, K, _p(W), _p(bias), N, act_in, act_out, _p(y),
                                     y.stride(0), _stream()))
    STATS.launches += 1


@_timed("timestep_embedding")
def timestep_embedding(t, freqs, cos_first, out):
    B = t.shape[0]
    half = freqs.shape[0]
    _lib.check(_lib.lib().nlc_timestep_embedding(_ctx(t), _p(t), B, _p(freqs), half, 1 if cos_first else 0, _p(out),
                                                 out.stride(0), _stream()))
    STATS.launches += 1


@_timed("row_norm")
def row_norm(x, out):
    B = x.shape[0]
    d = x[0].numel()
    _lib.check(_lib.lib().nlc_row_norm(_ctx(x), _p(x), B, d, _p(out), _stream()))
    STATS.launches += 1


@_timed("normalize_rows_")
def normalize_rows_(x):
    B = x.shape[0]
    d = x[0].numel()
    _lib.check(_lib.lib().nlc_normalize_rows(_ctx(x), _p(x), B, d, _stream()))
    STATS.launches += 1


@_timed("refine_sigma")
def refine_sigma(norms, B, d, sigma_in, norm_min, norm_max, refine, t_fixed, table, time_shift, sigma_out, t_out,
                 in_scale_out, slopes=None):
    _lib.check(_lib.lib().nlc_refine_sigma(
        _ctx(sigma_in), _p(norms), B, d, _p(sigma_in), sigma_in.numel(), norm_min, norm_max, 1 if refine else 0,
        float(t_fixed), _p(table), _p(slopes), table.numel() if table is not None else 0, int(time_shift),
        _p(sigma_out), _p(t_out), _p(in_scale_out), _stream()))
    STATS.launches += 1


@_timed("sigma_correct")
def sigma_correct(r, sigma, sigma_prev, update_prev, table, sigma_hat, sigma_prev_hat, t_hat, in_scale_out,
                  slopes=None):
    B = sigma.numel()
    _lib.check(_lib.lib().nlc_sigma_correct(
        _ctx(r), _p(r), _p(sigma), _p(sigma_prev), sigma_prev.numel(), B, 1 if update_prev else 0, _p(table),
        _p(slopes), table.numel(), _p(sigma_hat), _p(sigma_prev_hat), _p(t_hat), _p(in_scale_out), _stream()))
    STATS.launches += 1


@_timed("dynamic_threshold")
def dynamic_threshold_(x, ratio, max_value, s_out=None):
    """In place: x_b <- clamp(x_b, -s_b, s_b)/s_b with s_b = clamp(quantile(|x_b|, ratio), 1, max_value)."""
    B, d = x.shape[0], x[0].numel()
    _lib.check(_lib.lib().nlc_dynamic_threshold(_ctx(x), _p(x), B, d, float(ratio), float(max_value), _p(s_out),
                                                _stream()))
    STATS.launches += 1


@_timed("sigma_estimate")
def sigma_estimate(norms, last_norm, d, norm_max, sigma_prev_orig, sigma_prev, sigma_t, rates, table, slopes,
                   sigma_out, t_out):
    B = norms.numel()
    r4 = (C.c_float * 4)(*[float(v) for v in rates])
    _lib.check(_lib.lib().nlc_sigma_estimate(
        _ctx(norms), _p(norms), _p(last_norm), B, d, float(norm_max), float(sigma_prev_orig), _p(sigma_prev),
        sigma_prev.numel(), _p(sigma_t), sigma_t.numel(), C.byref(r4), _p(table), _p(slopes), table.numel(),
        _p(sigma_out), _p(t_out), _stream()))
    STATS.launches += 1


@_timed("pred_xstart")
def pred_xstart(xt, eps, sigma, clip, x0):
    B = xt.shape[0]
    d = xt[0].numel()
    _lib.check(_lib.lib().nlc_pred_xstart(_ctx(xt), _p(xt), _p(eps), _p(sigma), sigma.numel(), B, d, clip, _p(x0),
                                          _stream()))
    STATS.launches += 1


@_timed("pred_xprev")
def pred_xprev(sched, eta, x0, eps, xt, noise, learned_v, logvar_mode, min_var_coef, sigma, sigma_prev, x_prev,
               nan_flag=None):
    B = x0.shape[0]
    d = x0[0].numel()
    _lib.check(_lib.lib().nlc_pred_xprev(
        _ctx(x0), sched, float(eta), _p(x0), _p(eps), _p(xt), _p(noise), _p(learned_v), logvar_mode,
        float(min_var_coef), _p(sigma), sigma.numel(), _p(sigma_prev), sigma_prev.numel(), B, d, _p(x_prev),
        _p(nan_flag), _stream()))
    STATS.launches += 1


@_timed("best_update")
def best_update(loss_sum, count, best_val, flag, x0, best_x0):
    """Device-side `if mean(const) < best_val: best_x0 = x0.clone(); best_val = mean` (nlc_best_update)."""
    _lib.check(_lib.lib().nlc_best_update(_ctx(x0), _p(loss_sum), 1.0 / float(count), _p(best_val), _p(flag), _p(x0),
                                          _p(best_x0), x0.numel(), _stream()))
    STATS.launches += 2


# ---------------------------------------------------------------------------------------------- EDM sampler (fp64 state)
@_timed("edm_prepare")
def edm_prepare(x64, x32, sumsq_parts=None):
    B, d = x64.shape[0], x64[0].numel()
    _lib.check(_lib.lib().nlc_edm_prepare(_ctx(x64), _p(x64), B, d, _p(x32), _p(sumsq_parts), _stream()))
    STATS.launches += 1


@_timed("edm_eps")
def edm_eps(x64, x32, F, c_skip, c_out, div, eps, denoised=None, sumsq_parts=None):
    B, d = x64.shape[0], x64[0].numel()
    _lib.check(_lib.lib().nlc_edm_eps(_ctx(x64), _p(x64), _p(x32), _p(F), _p(c_skip), _p(c_out), _p(div), B, d, _p(eps),
                                      _p(denoised), _p(sumsq_parts), _stream()))
    STATS.launches += 1


@_timed("edm_mix")
def edm_mix(e1, den1, s1, e2, den2, s2, w1, w2, out, sums_parts=None):
    B, d = e1.shape[0], e1[0].numel()
    _lib.check(_lib.lib().nlc_edm_mix(_ctx(e1), _p(e1), _p(den1), _p(s1), _p(e2), _p(den2), _p(s2), float(w1), float(w2),
                                      B, d, _p(out), _p(sums_parts), _stream()))
    STATS.launches += 1


@_timed("edm_axpy")
def edm_axpy(x_hat, e, den, eps_scale, mul, coef, x_next):
    B, d = x_hat.shape[0], x_hat[0].numel()
    _lib.check(_lib.lib().nlc_edm_axpy(_ctx(x_hat), _p(x_hat), _p(e), _p(den), float(eps_scale or 0.0), _p(mul),
                                       _p(coef), B, d, _p(x_next), _stream()))
    STATS.launches += 1


# ---------------------------------------------------------------------- FID statistics (fid.py)
def fid_preprocess(x_nchw, from_pm1, quantize, resize, normalize, R_out, y_map, op_dtype):
    """[B,3,H,W] fp32 -> NHWC operand matrix of `y_map` (fid._Map): PNG round trip, bilinear 299, 2x - 1 (nlc_fid_preprocess)."""
    B, _, H, W = x_nchw.shape
    _lib.check(_lib.lib().nlc_fid_preprocess(_ctx(x_nchw), _p(x_nchw), B, H, W, int(from_pm1), int(quantize), int(resize),
                                             int(normalize), R_out, C.c_void_p(y_map.ptr), y_map.ld, op_dtype, _stream()))
    STATS.launches += 1


def im2col_nhwc(x_map, kh, kw, stride, pad, patches, op_dtype):
    """Feature map -> zero-padded patch matrix [M_pad, K_pad] (nlc_im2col_nhwc)."""
    _lib.check(_lib.lib().nlc_im2col_nhwc(_ctx(patches), C.c_void_p(x_map.ptr), op_dtype, x_map.ld, x_map.B, x_map.H, x_map.W,
                                          x_map.C, kh, kw, stride[0], stride[1], pad[0], pad[1], _p(patches), patches.shape[1],
                                          patches.shape[0], _stream()))
    STATS.launches += 1


def pool2d(x_map, stride, pad, mode, y_map, op_dtype):
    """3x3 max (mode 0) / average without padding count (mode 1) pooling between feature maps (nlc_pool2d)."""
    _lib.check(_lib.lib().nlc_pool2d(_ctx(x_map.t), C.c_void_p(x_map.ptr), op_dtype, x_map.ld, x_map.B, x_map.H, x_map.W,
                                     x_map.C, stride, pad, mode, C.c_void_p(y_map.ptr), y_map.ld, _stream()))
    STATS.launches += 1


def global_avgpool(x_map, y, op_dtype):
    _lib.check(_lib.lib().nlc_global_avgpool(_ctx(x_map.t), C.c_void_p(x_map.ptr), op_dtype, x_map.ld, x_map.B,
                                             x_map.H * x_map.W, x_map.C, _p(y), _stream()))
    STATS.launches += 1


def cov_accumulate(feats, sum64, outer64):
    """sum += sum_b f[b], outer += f^T f in fp64 (nlc_cov_accumulate)."""
    _lib.check(_lib.lib().nlc_cov_accumulate(_ctx(feats), _p(feats), feats.shape[0], feats.shape[1], _p(sum64), _p(outer64),
                                             _stream()))
    STATS.launches += 1
