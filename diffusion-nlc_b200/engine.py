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
from ._lib import NLC_BF16, NLC_F32
from .ops import Act

PRECISIONS = {"bf16": NLC_BF16, "tf32": NLC_F32}


class Feat:
    """An activation held as fp32 NHWC (residual stream: GroupNorm input, residual adds) and/or as a
    tensor-core operand copy (bf16, or tf32-rounded fp32)."""

    __slots__ = ("f32", "op")

    def __init__(self, f32=None, op=None):
        self.f32, self.op = f32, op

    @property
    def any(self):
        return self.f32 if self.f32 is not None else self.op

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
        self.chunk = 64 if self.op_dtype == NLC_BF16 else 32
        self._scratch = {}
        self._named = {}
        self._sizing = None

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
        return Act(self.scratch(tag, (B, H, W, C), torch.float32))

    def act_op(self, tag, B, H, W, C):
        return Act(self.scratch(tag, (B, H, W, C), self.op_torch))

    def bytes_allocated(self):
        tot = 0
        for t in list(self._scratch.values()) + list(self._named.values()):
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
def emit_groupnorm(pc, x32, gamma, beta, groups, eps, y_op, silu=True, scale=None, shift=None):
    ws = pc.gn_ws(x32.B, x32.H * x32.W, x32.C, groups)
    dt = pc.eng.op_dtype
    pc.add(lambda: ops.groupnorm(x32, groups, eps, gamma, beta, y_op, dt, ws(), silu=silu, scale=scale, shift=shift),
           "groupnorm %dx%dx%d B%d" % (x32.H, x32.W, x32.C, x32.B))


def emit_conv3x3(pc, src_op, w_packed, bias, Cout, dest, rowvec=None, resid=None, out_scale=1.0, stride=1, pad=1,
                 extra_src=None):
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
    pc.add(lambda: ops.conv_tc(srcs, segs, w_packed, Cout, B, Ho, Wo, dt, stride=stride, bias=bias, rowvec=rowvec,
                               resid=resid, out_scale=out_scale, out_f32=dest.f32, out_op=dest.op),
           "conv3x3 %dx%d %d->%d s%d B%d%s" % (Ho, Wo, src_op.C, Cout, stride, B, " +1x1" if extra_src is not None else ""))


def emit_conv1x1(pc, src_op, w_packed, bias, Cout, dest, resid=None, out_scale=1.0):
    dt = pc.eng.op_dtype
    B, H, W = src_op.B, src_op.H, src_op.W
    segs = [(0, 0, 0, 0, src_op.C)]
    pc.add(lambda: ops.conv_tc([src_op], segs, w_packed, Cout, B, H, W, dt, bias=bias, resid=resid,
                               out_scale=out_scale, out_f32=dest.f32, out_op=dest.op),
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
