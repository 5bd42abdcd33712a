"""Is the tensor-core convolution clock-limited by the board power cap?  Runs nlc_conv_tc on the dominant ADM / c2 shapes
for ~2 s per epilogue variant while sampling `nvidia-smi` (SM clock, board power), and prints TFLOP/s next to the median
clock and power: if the variants with the heavier epilogue run at a LOWER clock at the same (capped) power, the epilogue's
cost is energy (bytes moved, instructions issued), not exposed latency.
    python scripts/power_probe.py"""
import os
import subprocess
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nlc_b200 import ops
from nlc_b200._lib import NLC_BF16

dev = torch.device("cuda:0")


class Smi(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.rows, self.stop = [], False

    def run(self):
        while not self.stop:
            try:
                out = subprocess.run(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                a, b = [v.strip() for v in out.strip().split(",")[:2]]
                self.rows.append((int(a), float(b)))
            except Exception:
                pass
            time.sleep(0.1)


def run(B, H, Cin, Cout, name, seconds=2.0, **kw):
    x = ops.Act(torch.randn(B, H, H, Cin, device=dev).to(torch.bfloat16))
    w = ops.pack_conv_weight(torch.randn(Cout, Cin, 3, 3, device=dev) / (Cin * 9) ** 0.5, NLC_BF16)
    bias = torch.randn(Cout, device=dev) if kw.get("bias") else None
    resid = ops.Act(torch.randn(B, H, H, Cout, device=dev)) if kw.get("resid") else None
    st = ops.GnStats(torch.zeros(B * H * H // 32, Cout // 4, 2, device=dev)) if kw.get("stats") else None
    o32 = ops.Act(torch.empty(B, H, H, Cout, device=dev), 0, Cout, st) if kw.get("f32", True) else None
    o16 = ops.Act(torch.empty(B, H, H, Cout, device=dev, dtype=torch.bfloat16)) if kw.get("op") else None
    f = lambda: ops.conv_tc([x], ops.taps3x3(0, 0, Cin), w, Cout, B, H, H, NLC_BF16, bias=bias, resid=resid,
                            out_f32=o32, out_op=o16, stats=st is not None)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        f()
    e1.record()
    torch.cuda.synchronize()
    n = max(10, int(seconds * 1000 / (e0.elapsed_time(e1) / 5)))
    smi = Smi()
    smi.start()
    e0.record()
    for _ in range(n):
        f()
    e1.record()
    torch.cuda.synchronize()
    smi.stop = True
    ms = e0.elapsed_time(e1) / n
    fl = 2.0 * B * H * H * Cout * Cin * 9
    rows = smi.rows[len(smi.rows) // 3:] or [(0, 0.0)]
    mhz = sorted(r[0] for r in rows)[len(rows) // 2]
    watt = sorted(r[1] for r in rows)[len(rows) // 2]
    tf = fl / ms / 1e9
    # tensor-pipe utilisation at the clock actually held: 8192 dense 16-bit FLOP per clock per SM, 148 SMs
    print("%-30s %-26s %.3f ms %7.1f TFLOP/s  %4d MHz %5.0f W  pipe %.2f of the clock's peak" % (
        "%dx%d %d->%d B%d" % (H, H, Cin, Cout, B), name, ms, tf, mhz, watt, tf * 1e12 / (8192.0 * 148 * mhz * 1e6)), flush=True)


for shape in ((32, 256, 256, 256), (256, 64, 128, 128)):
    run(*shape, "op only", f32=False, op=True)
    run(*shape, "f32 only")
    run(*shape, "f32+bias+resid", bias=True, resid=True)
    run(*shape, "f32+op+bias+resid+stats", bias=True, resid=True, stats=True, op=True)
