"""Epilogue ablation of nlc_conv_tc on the two dominant conv shapes (c2: 64x64 128->128 B256; ADM: 256x256 256->256
B32): time with each epilogue feature switched on in turn."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nlc_b200 import ops
from nlc_b200._lib import NLC_BF16

dev = torch.device("cuda:0")


def run(B, H, Cin, Cout, name, **kw):
    x = ops.Act(torch.randn(B, H, H, Cin, device=dev).to(torch.bfloat16))
    w = ops.pack_conv_weight(torch.randn(Cout, Cin, 3, 3, device=dev) / (Cin * 9) ** 0.5, NLC_BF16)
    bias = torch.randn(Cout, device=dev) if kw.get("bias") else None
    rowvec = torch.randn(B, Cout, device=dev) if kw.get("rowvec") else None
    resid = ops.Act(torch.randn(B, H, H, Cout, device=dev)) if kw.get("resid") else None
    if kw.get("resid16"):  # 16-bit residual stream (nlc_conv_desc.resid_is_op)
        resid = ops.Act(torch.randn(B, H, H, Cout, device=dev).to(torch.bfloat16))
    st = ops.GnStats(torch.zeros(B * H * H // 32, Cout // 4, 2, device=dev)) if kw.get("stats") else None
    o32 = ops.Act(torch.empty(B, H, H, Cout, device=dev), 0, Cout, st) if kw.get("f32", True) else None
    o16 = ops.Act(torch.empty(B, H, H, Cout, device=dev, dtype=torch.bfloat16), 0, Cout, st if o32 is None else None) \
        if kw.get("op") else None
    f = lambda: ops.conv_tc([x], ops.taps3x3(0, 0, Cin), w, Cout, B, H, H, NLC_BF16, bias=bias, rowvec=rowvec, resid=resid,
                            out_f32=o32, out_op=o16, stats=st is not None)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        f()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    fl = 2.0 * B * H * H * Cout * Cin * 9
    print("%-34s %-28s %.3f ms %7.1f TFLOP/s" % ("%dx%d %d->%d B%d" % (H, H, Cin, Cout, B), name, ms, fl / ms / 1e9))


from nlc_b200 import _lib  # noqa: E402
if len(sys.argv) > 1:  # slab mode: 0 off, 1 Cout == 128, 2 every eligible layer
    _lib.check(_lib.lib().nlc_ctx_set(_lib.ctx(0), b"slab", int(sys.argv[1])))
    print("slab mode", sys.argv[1])
for shape in ((256, 64, 128, 128), (32, 256, 256, 256), (256, 32, 256, 256), (256, 64, 256, 128), (64, 64, 512, 512)):
    run(*shape, "op only", f32=False, op=True)
    run(*shape, "f32 only")
    run(*shape, "f32+bias+rowvec", bias=True, rowvec=True)
    run(*shape, "f32+bias+rowvec+stats", bias=True, rowvec=True, stats=True)
    run(*shape, "f32+bias+resid", bias=True, resid=True)
    run(*shape, "f32+op+bias+resid+stats", bias=True, resid=True, stats=True, op=True)
    run(*shape, "op+bias+rowvec+stats", f32=False, op=True, bias=True, rowvec=True, stats=True)
    run(*shape, "op+bias+resid16+stats", f32=False, op=True, bias=True, resid16=True, stats=True)
