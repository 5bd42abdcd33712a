"""Launches for an ncu --set full capture of the epilogue-bound GEMM shapes of a c2 step (batch 256, fp16, TMA epilogue):
the im2col input convolution (64x64, K 64, N 128), the q/k/v and proj 1x1 GEMMs of the 16x16 attention blocks (K 256, N 768 /
N 256 + 16-bit residual), the 64x64 128->128 slab layer with the ResBlock epilogue, and the 4x4 K 4608 N 512 fp32-output layer.
    ncu --set full --clock-control none --import-source on -k regex:conv_ -f -o gpurun_out/X python scripts/epi_ncu2.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nlc_b200 import ops
from nlc_b200._lib import NLC_F16

dev = torch.device("cuda:0")
B = 256
h16 = lambda *s: torch.randn(*s, device=dev).to(torch.float16)


def gemm(H, Cin, Cout, k, resid, stats, f32=False, reps=2):
    x = ops.Act(h16(B, H, H, Cin))
    w = ops.pack_conv_weight(torch.randn(Cout, Cin, k, k, device=dev) / (Cin * k * k) ** 0.5, NLC_F16)
    bias = torch.randn(Cout, device=dev)
    st = ops.GnStats(torch.zeros(B * H * H // 32, Cout // 4, 2, device=dev)) if stats else None
    r = ops.Act(h16(B, H, H, Cout)) if resid else None
    if f32:
        r = ops.Act(torch.randn(B, H, H, Cout, device=dev)) if resid else None
        o32 = ops.Act(torch.empty(B, H, H, Cout, device=dev), 0, Cout, st)
        o16 = None
    else:
        o32 = None
        o16 = ops.Act(torch.empty(B, H, H, Cout, device=dev, dtype=torch.float16), 0, Cout, st)
    segs = ops.taps3x3(0, 0, Cin) if k == 3 else [(0, 0, 0, 0, Cin)]
    for _ in range(reps):
        ops.conv_tc([x], segs, w, Cout, B, H, H, NLC_F16, bias=bias, resid=r, out_f32=o32, out_op=o16, stats=stats)
    torch.cuda.synchronize()


gemm(64, 64, 128, 1, False, True)      # launches 0-1: input convolution as a GEMM over the im2col matrix
gemm(16, 256, 768, 1, False, False)    # 2-3: q/k/v
gemm(16, 256, 256, 1, True, True)      # 4-5: proj + residual
gemm(64, 128, 128, 3, True, True)      # 6-7: slab kernel, ResBlock conv2
gemm(4, 512, 512, 3, True, False, f32=True)  # 8-9: 4x4 level, fp32 residual stream
