"""How closely does the REFERENCE's own GPU path reproduce the reference's CPU run?  (test-infrastructure script: runs
the oracle, which is pinned torch.equal to the reference, through PyTorch eager / cuDNN on the GPU -
tests/parity_util.oracle_c2_loop_on_gpu, also the control of tests/test_gpu_bench_arch.py.)

The north-star gates the 16-bit mode at "final-image PSNR >= 45 dB against the reference".  The reference run that the
golden fixture tests/golden/loop_c2_100.pt records is its fp32 CPU run; the reference's DEFAULT GPU run uses TF32
convolutions (torch.backends.cudnn.allow_tf32 = True).  This script runs the same c2 100-step NLC loop (batch 4, same
seeded weights, same noise draws) through the oracle on the GPU in three arithmetic settings and reports, against the
golden final image: PSNR free-running, PSNR with the golden run's time buckets forced (t = searchsorted(sigma) taken from
the recorded run, everything else free), and how many of the 4 samples crossed a time-bucket edge.

    gpurun -- python scripts/ref_gpu_selfparity.py > gpurun_out/r02_ref_gpu_selfparity.log
"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from parity_util import oracle_c2_loop_on_gpu  # noqa: E402


def psnr(a, b):
    return 10 * math.log10(4.0 / max(torch.mean((a.double().cpu() - b.double().cpu()) ** 2).item(), 1e-30))


def main():
    g = torch.load(os.path.join(ROOT, "tests", "golden", "loop_c2_100.pt"), weights_only=True)
    shape = (4, 3, 64, 64)
    torch.manual_seed(5)
    z = torch.randn(shape)
    noises = [torch.randn(shape) for _ in range(100)]
    settings = [("fp32 cuDNN (allow_tf32 off)", False, None),
                ("TF32 convolutions (the reference's default GPU run)", True, None),
                ("torch.autocast(float16)", True, torch.float16)]
    for label, tf32, autocast in settings:
        res = {}
        for forced in (True, False):
            out, flips = oracle_c2_loop_on_gpu(g, z, noises, tf32=tf32, autocast=autocast, forced=forced)
            res[forced] = (psnr(out, g["final"]), flips)
        print("%-55s PSNR %.1f dB with the CPU run's time buckets, %.1f dB entirely free (%d of 4 samples crossed a time "
              "bucket)" % (label, res[True][0], res[False][0], res[False][1]), flush=True)


if __name__ == "__main__":
    main()
