"""Python-side wrappers of the C ABI: thin argument marshalling from torch tensors (device memory +
current stream) to libnlc_b200 entry points.  No arithmetic happens here."""
import ctypes as C

import torch

from . import _lib
from ._lib import NLC_BF16, NLC_F32

OP_DTYPES = {NLC_BF16: torch.bfloat16, NLC_F32: torch.float32}


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

    __slots__ = ("t", "c0", "C")

    def __init__(self, t, c0=0, C=None):
        assert t.dim() == 4 and t.is_contiguous(), "Act wants a contiguous [B,H,W,C] tensor"
        self.t, self.c0 = t, c0
        self.C = t.shape[3] - c0 if C is None else C

    B = property(lambda s: s.t.shape[0])
    H = property(lambda s: s.t.shape[1])
    W = property(lambda s: s.t.shape[2])
    ld = property(lambda s: s.t.shape[3])
    dtype = property(lambda s: s.t.dtype)
    ptr = property(lambda s: s.t.data_ptr() + s.c0 * s.t.element_size())

    def slice(self, c0, C):
        return Act(self.t, self.c0 + c0, C)

    def dense(self):
        """[B,H,W,C] torch view (for tests)."""
        return self.t[..., self.c0:self.c0 + self.C]


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
    if op_dtype == NLC_BF16:
        return k.to(torch.bfloat16).contiguous()
    return round_tf32_(k.clone())


def taps3x3(src, c0, nch, pad=1):
    """Nine K segments of a 3x3 convolution over channels [c0,c0+nch) of source `src`."""
    return [(src, kh - pad, kw - pad, c0, nch) for kh in range(3) for kw in range(3)]


def conv_tc(srcs, segs, weight, Cout, B, Ho, Wo, op_dtype, stride=1, bias=None, rowvec=None, resid=None,
            out_scale=1.0, out_f32=None, out_op=None):
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
    d.out_scale = out_scale
    if out_f32 is not None:
        d.out_f32, d.ld_out_f32 = out_f32.ptr, out_f32.ld
    if out_op is not None:
        d.out_op, d.ld_out_op = out_op.ptr, out_op.ld
    _lib.check(_lib.lib().nlc_conv_tc(_ctx(srcs[0].t), C.byref(d), _stream()))
