"""Execution engine shared by the three UNet families: device buffers, the operand precision mode and the
block emitters (ResNet block, attention block, resampling convs) expressed as calls into libnlc_b200.

A network is "planned" once per batch size: every activation gets a fixed device buffer (scratch buffers are
shared between layers, skip tensors are written straight into the channel slice of the concat buffer that the
decoder block will read), and the plan is a flat list of closures.  Running a plan only enqueues kernels on the
current stream, so it can be replayed or captured in a CUDA graph.
"""
import os

import torch

from . import ops
from ._lib import NLC_BF16, NLC_F16, NLC_F32, NLC_F32X3
from .ops import Act, GnStats

# bf16 / fp16: throughput modes (same MMA rate; fp16 keeps 3 more mantissa bits and is the reference's own reduced
# precision, src/fp16_util.py).  tf32: one kind::tf32 MMA on tf32-rounded fp32 operands (what cuDNN does for the reference's
# default GPU run).  fp32: the accuracy mode, unrounded fp32 operands and weights, 3 x tf32 split products.
PRECISIONS = {"bf16": NLC_BF16, "fp16": NLC_F16, "tf32": NLC_F32, "fp32": NLC_F32X3}


class Feat:
    """An activation held as fp32 NHWC (residual stream: GroupNorm input, residual adds) and/or as a
    tensor-core operand copy (bf16, or tf32-rounded fp32)."""

    __slots__ = ("f32", "op")

    def __init__(self, f32=None, op=None):
        self.f32, self.op = f32, op

    @property
    def any(self):
        return self.f32 if self.f32 is not None else self.op

    # what a GroupNorm or a residual add reads: the fp32 tensor, or - 16-bit-activation plans (Engine.r16) - the operand copy
    res = any

    B = property(lambda s: s.any.B)
    H = property(lambda s: s.any.H)
    W = property(lambda s: s.any.W)
    C = property(lambda s: s.any.C)


class Engine:
    def __init__(self, device, precision="bf16"):
        if precision not in PRECISIONS:
            raise ValueError("precision must be one of %s" % list(PRECISIONS))
        self.device = torch.device(device)
        self.precision = precision
        self.op_dtype = PRECISIONS[precision]
        self.op_torch = ops.OP_DTYPES[self.op_dtype]
        self.chunk = 64 if self.op_dtype in (NLC_BF16, NLC_F16) else 32
        self._scratch = {}
        self._named = {}
        self._named_stats = {}
        self._sizing = None
        # GroupNorm statistics from the producing conv's epilogue (NLC_FUSE_GN_STATS=0 restores the separate pass)
        self.fuse_gn_stats = os.environ.get("NLC_FUSE_GN_STATS", "1") != "0"
        # 16-bit tensor between a ResBlock's two convolutions (act_h): on in the fp16 mode, whose rounding (2^-11) is far
        # below the mode's own error budget; off in bf16 (2^-8 on the GroupNorm input costs ~1 dB of the 45 dB gate);
        # NLC_H16=0|1 overrides
        h16 = os.environ.get("NLC_H16")
        self.h16 = (self.op_dtype == NLC_F16) if h16 is None else (h16 != "0" and self.op_dtype in (NLC_BF16, NLC_F16))
        # 16-bit RESIDUAL STREAM (`stream_feat`): at the levels whose GroupNorm statistics come from the conv epilogues
        # (>= 128 pixels per image) every activation - block outputs, skip / concat buffers, the residual operand of the
        # second conv - exists in the operand dtype only.  The convolutions are bound by bytes and by the board power cap, not
        # by the tensor pipe (profiles/r02h_power_probe.log): a ResBlock moves 18 instead of 26-32 bytes per element.  On in
        # the fp16 mode (2^-11 per rounding, the same the operand copies already carry); NLC_R16=0|1 overrides.
        r16 = os.environ.get("NLC_R16")
        self.r16 = self.h16 and self.fuse_gn_stats and ((self.op_dtype == NLC_F16) if r16 is None else r16 != "0")

    # ------------------------------------------------------------------ buffers
    def plan_two_pass(self, build):
        """Run `build()` twice: a sizing pass on the meta device that only records the largest request per scratch
        tag, then the real pass against scratch buffers allocated once at their final size (a scratch buffer that
        grew between two layers of one plan would otherwise stay alive twice)."""
        self._sizing = {}
        try:
            build()
            sizes = self._sizing
        finally:
            self._sizing = None
        for (tag, dtype), n in sizes.items():
            t = self._scratch.get((tag, dtype))
            if t is None or t.numel() < n:
                self._scratch[(tag, dtype)] = torch.empty(n, device=self.device, dtype=dtype)
        return build()

    def named(self, name, shape, dtype):
        """A buffer that lives as long as the plan (skip/concat tensors, outputs)."""
        if self._sizing is not None:
            return torch.empty(shape, device="meta", dtype=dtype)
        key = (name, tuple(shape), dtype)
        t = self._named.get(key)
        if t is None:
            t = torch.empty(shape, device=self.device, dtype=dtype)
            self._named[key] = t
        return t

    def scratch(self, tag, shape, dtype):
        """A scratch buffer shared by every layer that asks for the same tag (stream order makes reuse safe).
        Grows to the largest request; returns a contiguous view of the requested shape."""
        n = 1
        for s in shape:
            n *= s
        key = (tag, dtype)
        if self._sizing is not None:
            self._sizing[key] = max(self._sizing.get(key, 0), n)
            return torch.empty(shape, device="meta", dtype=dtype)
        t = self._scratch.get(key)
        if t is None or t.numel() < n:
            t = torch.empty(n, device=self.device, dtype=dtype)
            self._scratch[key] = t
        return t[:n].view(*shape)

    def act_f32(self, tag, B, H, W, C):
        """fp32 scratch activation; when a conv epilogue can write GroupNorm partials for it, a (fresh) stats holder
        over the tag's stats scratch comes with it."""
        a = Act(self.scratch(tag, (B, H, W, C), torch.float32))
        if self.fuse_gn_stats and GnStats.eligible(B, H, W, C):
            st = self.scratch(tag + ".stats", (B * H * W // 32, C // 4, 2), torch.float32)
            if self._sizing is None:
                a.stats = GnStats(st)
        return a

    def act_h(self, tag, B, H, W, C):
        """The activation between a ResBlock's two convolutions: it only feeds a GroupNorm.  In the 16-bit modes with the
        statistics taken from the producing conv's fp32 accumulators it is kept in the OPERAND dtype only (`h16`): the conv
        epilogue writes 2 instead of 4 bytes per element and the GroupNorm apply pass reads 2 instead of 4 - the kernels are
        bound by bytes through L2 / HBM, not by the tensor pipe.  Otherwise fp32 as before."""
        if self.h16 and self.fuse_gn_stats and GnStats.eligible(B, H, W, C):
            a = Act(self.scratch(tag + ".16", (B, H, W, C), self.op_torch))
            st = self.scratch(tag + ".stats", (B * H * W // 32, C // 4, 2), torch.float32)
            if self._sizing is None:
                a.stats = GnStats(st)
            return a
        return self.act_f32(tag, B, H, W, C)

    def level16(self, B, H, W, C):
        """True when the activations of this shape live in the operand dtype only (16-bit residual stream)."""
        return self.r16 and GnStats.eligible(B, H, W, C)

    def stream_feat(self, tag, B, H, W, C, need_op=False):
        """A block output on the residual stream: fp32 scratch (+ an operand copy when a conv reads it raw), or - 16-bit
        residual stream - one operand-dtype scratch with its GroupNorm statistics holder."""
        if self.level16(B, H, W, C):
            a = Act(self.scratch(tag + ".16", (B, H, W, C), self.op_torch))
            st = self.scratch(tag + ".stats", (B * H * W // 32, C // 4, 2), torch.float32)
            if self._sizing is None:
                a.stats = GnStats(st)
            return Feat(op=a)
        return Feat(f32=self.act_f32(tag, B, H, W, C), op=self.act_op(tag + ".op", B, H, W, C) if need_op else None)

    def cat_buffers(self, k, B, H, W, C):
        """The concat buffer of skip k: (fp32, operand) named tensors; the fp32 one is None on the 16-bit residual stream."""
        c16 = self.named("cat16.%d" % k, (B, H, W, C), self.op_torch)
        if self.level16(B, H, W, C):
            return None, c16
        return self.named("cat32.%d" % k, (B, H, W, C), torch.float32), c16

    def cat_view(self, c32, c16, c0=0, C=None):
        """Feat over channels [c0, c0+C) of a concat buffer pair (statistics holder on the tensor a GroupNorm will read)."""
        if c32 is None:
            return Feat(op=self.with_stats(Act(c16, c0, C)))
        return Feat(self.with_stats(Act(c32, c0, C)), Act(c16, c0, C))

    def with_stats(self, act):
        """Attach the stats holder of a plan-lifetime (named) buffer to a view of it."""
        t = act.t
        if self._sizing is not None or not self.fuse_gn_stats:
            return act
        B, H, W, C = t.shape
        if not GnStats.eligible(B, H, W, C):
            return act
        key = t.data_ptr()
        st = self._named_stats.get(key)
        if st is None:
            st = GnStats(torch.empty((B * H * W // 32, C // 4, 2), device=self.device, dtype=torch.float32))
            self._named_stats[key] = st
        act.stats = st
        return act

    def act_op(self, tag, B, H, W, C):
        return Act(self.scratch(tag, (B, H, W, C), self.op_torch))

    def bytes_allocated(self):
        tot = 0
        for t in list(self._scratch.values()) + list(self._named.values()) + [s.t for s in self._named_stats.values()]:
            tot += t.numel() * t.element_size()
        return tot

    # ------------------------------------------------------------------ weights
    def pack3x3(self, w, extra=None):
        return ops.pack_conv_weight(w.to(self.device), self.op_dtype,
                                    extra.to(self.device) if extra is not None else None)

    def dev32(self, t):
        return t.detach().to(self.device, torch.float32).contiguous()


class PlanCtx:
    """Scratch state while emitting one plan: the step list plus lazily sized shared workspaces."""

    def __init__(self, eng, B):
        self.eng, self.B = eng, B
        self.steps = []
        self._gn_ws_floats = 0
        self._attn_ws_bytes = 0

    def add(self, fn, label=None):
        if label is not None:
            fn.label = label
        self.steps.append(fn)

    def gn_ws(self, B, HW, C, groups):
        self._gn_ws_floats = max(self._gn_ws_floats, ops.groupnorm_ws(B, HW, C, groups))
        eng = self.eng
        return lambda: eng.scratch("gn_ws", (self._gn_ws_floats,), torch.float32)

    def attn_ws(self, B, T, heads, dh):
        self._attn_ws_bytes = max(self._attn_ws_bytes, ops.attention_ws(self.eng.op_dtype, B, T, heads, dh))
        eng = self.eng
        return lambda: eng.scratch("attn_ws", (max(self._attn_ws_bytes, 16),), torch.uint8)


# ---------------------------------------------------------------------- block emitters
def emit_groupnorm(pc, x32, gamma, beta, groups, eps, y_op, silu=True, scale=None, shift=None, resample=0):
    """resample 1 / 2: y_op is the activated tensor nearest-x2 upsampled / 2x2 average pooled."""
    ws = pc.gn_ws(x32.B, x32.H * x32.W, x32.C, groups)
    dt = pc.eng.op_dtype
    fused = x32.stats is not None and x32.stats.covers(x32.c0, x32.C)
    assert pc.eng._sizing is not None or x32.dtype == torch.float32 or fused, \
        "a 16-bit GroupNorm input needs fused statistics"
    pc.add(lambda: ops.groupnorm(x32, groups, eps, gamma, beta, y_op, dt, ws(), silu=silu, scale=scale, shift=shift,
                                 use_stats=fused, resample=resample),
           "groupnorm %dx%dx%d B%d%s%s%s" % (x32.H, x32.W, x32.C, x32.B, " fused-stats" if fused else "",
                                            (" up2", " pool2")[resample - 1] if resample else "",
                                            " x16" if x32.dtype != torch.float32 else ""))


def h_feat(h):
    """Feat around an `Engine.act_h` activation (operand dtype only, or fp32 only)."""
    return Feat(f32=h) if h.dtype == torch.float32 else Feat(op=h)


def _want_stats(dest, Cout):
    """True when the conv writing `dest` should also write GroupNorm partials (and records the coverage)."""
    a = dest.f32 if (dest.f32 is not None and (dest.f32.stats is not None or dest.op is None)) else dest.op
    if a is None or a.stats is None or Cout % 4 != 0 or a.c0 % 4 != 0:
        return False
    a.stats.covered.append((a.c0, a.c0 + Cout))
    return True


def emit_conv3x3(pc, src_op, w_packed, bias, Cout, dest, rowvec=None, resid=None, out_scale=1.0, stride=1, pad=1,
                 extra_src=None, resid_mode=0):
    """3x3 conv over operand `src_op` (+ optional fused 1x1 over `extra_src`, already packed behind the 3x3
    weights) into Feat `dest` (fp32 and/or operand copy)."""
    dt = pc.eng.op_dtype
    B = src_op.B
    Ho, Wo = (src_op.H // stride, src_op.W // stride)
    srcs = [src_op]
    segs = ops.taps3x3(0, 0, src_op.C, pad=pad)
    if extra_src is not None:
        srcs.append(extra_src)
        segs = segs + [(1, 0, 0, 0, extra_src.C)]
    st = _want_stats(dest, Cout)
    pc.add(lambda: ops.conv_tc(srcs, segs, w_packed, Cout, B, Ho, Wo, dt, stride=stride, bias=bias, rowvec=rowvec,
                               resid=resid, out_scale=out_scale, out_f32=dest.f32, out_op=dest.op, stats=st,
                               resid_mode=resid_mode),
           "conv3x3 %dx%d %d->%d s%d B%d%s" % (Ho, Wo, src_op.C, Cout, stride, B, " +1x1" if extra_src is not None else ""))


def emit_upsample_conv3x3(pc, src_op, phase_w, bias, Cout, dest, rowvec=None):
    """Upsample(with_conv) = nearest x2 followed by a 3x3 conv (src/unet_ddim.py:58-74), computed on the LOW-resolution
    operand `src_op` as four sub-pixel phase convolutions with 2x2 taps each (ops.upsample_phase_weights) that write
    their quarter of `dest` (2H x 2W) in place (nlc_conv_desc.out_up): 16 instead of 36 multiplies per output and no
    replicated operand in HBM."""
    dt = pc.eng.op_dtype
    B, H, W = src_op.B, src_op.H, src_op.W
    st = _want_stats(dest, Cout)
    for a in (0, 1):
        for b in (0, 1):
            segs = ops.upsample_phase_taps(0, 0, src_op.C, a, b)
            pc.add(lambda segs=segs, w=phase_w[(a, b)], ab=(a, b): ops.conv_tc(
                [src_op], segs, w, Cout, B, H, W, dt, bias=bias, rowvec=rowvec, out_f32=dest.f32, out_op=dest.op, stats=st,
                out_up=ab),
                "upconv3x3 phase %d%d %dx%d %d->%d B%d" % (a, b, H, W, src_op.C, Cout, B))


def upsample_conv_eligible(H, W):
    """The low-resolution conv must tile inside one image and write whole GroupNorm partial blocks (NLC_UPCONV=0 keeps
    the replicated-operand path for A/B measurements)."""
    import os
    return (H * W >= 128 and (H & (H - 1)) == 0 and (W & (W - 1)) == 0 and os.environ.get("NLC_UPCONV", "1") != "0")


def emit_conv_in(pc, x_nchw, in_scale_fn, w_packed, bias, Cout, dest, w_f32=None):
    """The network's input convolution.  bf16 mode: im2col of the NCHW image (per-sample input scale folded in)
    followed by a K = 64 tensor-core GEMM.  tf32 (accuracy) mode: the fp32 CUDA-core kernel, so that the first layer
    stays exact as in the reference (rounding the image itself to tf32 costs sigma_hat ~1e-4 and with it the
    occasional time-bucket flip, tests/test_gpu_sampler.py).  `in_scale_fn()` returns the [B] scale or None."""
    eng = pc.eng
    dt = eng.op_dtype
    if dt not in (NLC_BF16, NLC_F16) and w_f32 is not None:
        pc.add(lambda: ops.conv_in_nchw(x_nchw, in_scale_fn(), w_f32, bias, dest.f32, dest.op, dt), "conv_in (fp32)")
        return
    B, _, H, W = x_nchw.shape
    kp = eng.chunk
    patches = eng.act_op("im2col", B, H, W, kp)
    pc.add(lambda: ops.im2col_in(x_nchw, in_scale_fn(), patches, dt), "im2col_in %dx%d B%d" % (H, W, B))
    st = _want_stats(dest, Cout)
    pc.add(lambda: ops.conv_tc([patches], [(0, 0, 0, 0, kp)], w_packed, Cout, B, H, W, dt, bias=bias, out_f32=dest.f32,
                               out_op=dest.op, stats=st), "conv_in %dx%d ->%d B%d" % (H, W, Cout, B))


def emit_conv_out(pc, src_op, w_f32, bias_f32, packed, out_nchw):
    """The network's output convolution (C -> 3 | 6 channels, NCHW fp32 for the sampler).  16-bit modes: one 64-channel
    tensor-core tile (weights zero-padded, `packed` = ops.pack_conv_out_weight(...)) into an fp32 scratch, then the
    real channels are copied out as NCHW; the fp32-container (accuracy) modes keep the fp32 CUDA-core kernel."""
    eng = pc.eng
    dt = eng.op_dtype
    Cout = w_f32.shape[0]
    if dt not in (NLC_BF16, NLC_F16) or packed is None or Cout > 8:
        pc.add(lambda: ops.conv_out_nchw(src_op, dt, w_f32, bias_f32, out_nchw), "conv_out (fp32 direct)")
        return
    wp, bp = packed
    B, H, W = src_op.B, src_op.H, src_op.W
    tmp = Act(eng.scratch("conv_out.tmp", (B, H, W, wp.shape[0]), torch.float32))
    segs = ops.taps3x3(0, 0, src_op.C)
    pc.add(lambda: ops.conv_tc([src_op], segs, wp, wp.shape[0], B, H, W, dt, bias=bp, out_f32=tmp),
           "conv_out %dx%d %d->%d (padded to %d) B%d" % (H, W, src_op.C, Cout, wp.shape[0], B))
    pc.add(lambda: ops.nhwc_head_to_nchw(tmp, Cout, out_nchw), "conv_out -> NCHW")


def emit_conv1x1(pc, src_op, w_packed, bias, Cout, dest, resid=None, out_scale=1.0):
    dt = pc.eng.op_dtype
    B, H, W = src_op.B, src_op.H, src_op.W
    segs = [(0, 0, 0, 0, src_op.C)]
    st = _want_stats(dest, Cout)
    pc.add(lambda: ops.conv_tc([src_op], segs, w_packed, Cout, B, H, W, dt, bias=bias, resid=resid,
                               out_scale=out_scale, out_f32=dest.f32, out_op=dest.op, stats=st),
           "conv1x1 %dx%d %d->%d B%d" % (H, W, src_op.C, Cout, B))


def emit_attention(pc, qkv_op, q_off, k_off, v_off, head_stride, heads, dh, scale, out_op):
    ws = pc.attn_ws(qkv_op.B, qkv_op.H * qkv_op.W, heads, dh)
    dt = pc.eng.op_dtype
    pc.add(lambda: ops.attention(qkv_op, dt, q_off, k_off, v_off, head_stride, heads, dh, scale, out_op, ws()),
           "attention T%d heads%d dh%d B%d" % (qkv_op.H * qkv_op.W, heads, dh, qkv_op.B))


def run(steps):
    if os.environ.get("NLC_SYNC") == "1":  # debugging aid: localise an asynchronous kernel fault to its plan step
        for i, fn in enumerate(steps):
            fn()
            try:
                torch.cuda.synchronize()
            except Exception as e:
                raise RuntimeError("plan step %d (%s) faulted: %s" % (i, getattr(fn, "label", "?"), e)) from e
        return
    for fn in steps:
        fn()
