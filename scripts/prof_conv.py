"""Short driver for ncu: a handful of launches of the two conv shapes that dominate the c2 step."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nlc_b200 import ops
from nlc_b200._lib import NLC_BF16

dev = torch.device("cuda:0")
for (B, H, Cin, Cout) in ((256, 64, 128, 128), (256, 32, 256, 256)):
    x = ops.Act(torch.randn(B, H, H, Cin, device=dev).to(torch.bfloat16))
    w = ops.pack_conv_weight(torch.randn(Cout, Cin, 3, 3, device=dev) / (Cin * 9) ** 0.5, NLC_BF16)
    b = torch.randn(Cout, device=dev)
    o32 = ops.Act(torch.empty(B, H, H, Cout, device=dev))
    o16 = ops.Act(torch.empty(B, H, H, Cout, device=dev, dtype=torch.bfloat16))
    for _ in range(3):
        ops.conv_tc([x], ops.taps3x3(0, 0, Cin), w, Cout, B, H, H, NLC_BF16, bias=b, out_f32=o32, out_op=o16)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.conv_tc([x], ops.taps3x3(0, 0, Cin), w, Cout, B, H, H, NLC_BF16, bias=b, out_f32=o32, out_op=o16)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    fl = 2.0 * B * H * H * Cout * Cin * 9
    print("conv %dx%d %d->%d B%d: %.3f ms %.1f TFLOP/s" % (H, H, Cin, Cout, B, ms, fl / ms / 1e9))
