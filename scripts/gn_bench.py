"""GroupNorm apply micro-benchmark (fused-statistics path): GB/s against the measured HBM copy peak.
    python scripts/gn_bench.py [B H W C]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nlc_b200 import ops
from nlc_b200._lib import NLC_BF16

dev = torch.device("cuda:0")
shapes = [tuple(int(v) for v in sys.argv[1:5])] if len(sys.argv) >= 5 else [(32, 256, 256, 256), (32, 128, 128, 256),
                                                                           (256, 64, 64, 128), (256, 32, 32, 256)]
for B, H, W, C, in16 in [s + (i,) for s in shapes for i in (False, True)]:
    xt = torch.randn(B, H, W, C, device=dev)
    x = ops.Act(xt.to(torch.bfloat16) if in16 else xt, 0, C, ops.GnStats(torch.rand(B * H * W // 32, C // 4, 2, device=dev)))
    x.stats.covered.append((0, C))
    y = ops.Act(torch.empty(B, H, W, C, device=dev, dtype=torch.bfloat16))
    gam, bet = torch.randn(C, device=dev), torch.randn(C, device=dev)
    ws = torch.zeros(ops.groupnorm_ws(B, H * W, C, 32), device=dev)
    f = lambda: ops.groupnorm(x, 32, 1e-5, gam, bet, y, NLC_BF16, ws, silu=True, use_stats=True)
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
    gb = B * H * W * C * (4 if in16 else 6) / 1e9
    print("GN apply (+finalize) B%d %dx%d C%d %s input: %.3f ms  %.0f GB/s" % (B, H, W, C, "16-bit" if in16 else "fp32", ms,
                                                                              gb / ms * 1e3))
