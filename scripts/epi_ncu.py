"""Two launches of nlc_conv_tc for an ncu --set full capture: the ADM 256x256 256->256 layer (batch 32) with the light
epilogue (operand output only) and with the 16-bit-stream ResBlock epilogue (operand output + bias + per-sample row + GroupNorm
partials), then the same pair for the c2 64x64 128->128 layer (batch 256).
    ncu --set full --clock-control none --import-source on -k regex:conv_ -f -o gpurun_out/X python scripts/epi_ncu.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nlc_b200 import ops
from nlc_b200._lib import NLC_F16

dev = torch.device("cuda:0")
for B, H, Cin, Cout in ((32, 256, 256, 256), (256, 64, 128, 128)):
    x = ops.Act(torch.randn(B, H, H, Cin, device=dev).to(torch.float16))
    w = ops.pack_conv_weight(torch.randn(Cout, Cin, 3, 3, device=dev) / (Cin * 9) ** 0.5, NLC_F16)
    bias, rowvec = torch.randn(Cout, device=dev), torch.randn(B, Cout, device=dev)
    resid = ops.Act(torch.randn(B, H, H, Cout, device=dev).to(torch.float16))
    st = ops.GnStats(torch.zeros(B * H * H // 32, Cout // 4, 2, device=dev))
    o_plain = ops.Act(torch.empty(B, H, H, Cout, device=dev, dtype=torch.float16))
    o_stats = ops.Act(torch.empty(B, H, H, Cout, device=dev, dtype=torch.float16), 0, Cout, st)
    segs = ops.taps3x3(0, 0, Cin)
    ops.conv_tc([x], segs, w, Cout, B, H, H, NLC_F16, out_op=o_plain)
    ops.conv_tc([x], segs, w, Cout, B, H, H, NLC_F16, bias=bias, rowvec=rowvec, out_op=o_stats, stats=True)
    ops.conv_tc([x], segs, w, Cout, B, H, H, NLC_F16, bias=bias, resid=resid, out_op=o_stats, stats=True)
    torch.cuda.synchronize()
