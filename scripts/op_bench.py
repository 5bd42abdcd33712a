"""Operator and DDNM-step kernels against the HBM roofline at the benchmark size
R = 256, batch 64: CUDA-event time per call (20 calls after 5 warm-ups, rotating over four sets of input / output tensors
so that consecutive calls cannot hit in the 126 MB L2), algorithmic bytes = the tensors a fused implementation has to read and write once, achieved
GB/s against MEASURED_PEAKS.json's copy bandwidth.  Calls are replayed from a CUDA graph (see timeit).

    python scripts/op_bench.py [B] [R]
"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nlc_b200  # noqa: E402,F401
from nlc_b200 import svd_operators as P  # noqa: E402
from nlc_b200.constraint_functions import _gauss_kernel  # noqa: E402


def timeit(fn, n=20, warm=5):
    """Time per call on the device.  The calls are captured once into a CUDA graph (n calls per replay) so that the
    host-side cost of the Python wrapper (ctypes, torch.empty) does not hide behind or pad the kernels: these kernels
    run for 20-100 us, the same order as one eager call's host time."""
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        with torch.cuda.graph(graph, stream=side):
            for _ in range(n):
                fn()
    torch.cuda.synchronize()
    graph.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    graph.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    R = int(sys.argv[2]) if len(sys.argv) > 2 else 256
    dev = torch.device("cuda:0")
    peak = 6527.8
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    C = 3
    d = C * R * R
    gen = torch.Generator().manual_seed(1)
    mask = torch.ones(R, R)
    mask[R // 4:3 * R // 4, R // 4:3 * R // 4] = 0
    mr = torch.nonzero(mask.reshape(-1) == 0).long().reshape(-1) * 3
    missing = torch.cat([mr, mr + 1, mr + 2])
    ops = {
        "colorization": P.Colorization(R, dev),
        "sr_averagepooling x4": P.SuperResolution(C, R, 4, dev),
        "inpainting box": P.Inpainting(C, R, missing, dev),
        "cs_walshhadamard x4": P.WalshHadamardCS(C, R, 4, torch.randperm(R * R, generator=gen), dev),
        "deblur_gauss": P.Deblurring(_gauss_kernel(5, 10), C, R, dev),
        "denoising": P.Denoising(C, R, dev),
    }
    NBUF = 4  # rotate over four input / output sets (>= 0.5 GB per round) so that nothing is served from the 126 MB L2
    xts = [torch.randn(B, C, R, R, device=dev) for _ in range(NBUF)]
    ets = [torch.randn(B, 2 * C, R, R, device=dev) for _ in range(NBUF)]
    zs = [torch.randn(B, C, R, R, device=dev) for _ in range(NBUF)]
    print("B=%d R=%d  HBM peak %.0f GB/s (MEASURED_PEAKS.json)" % (B, R, peak))
    print("%-22s %-12s %9s %9s %8s %6s" % ("operator", "call", "ms", "alg MB", "GB/s", "frac"))
    for name, op in ops.items():
        ys = [op.A(xt.reshape(B, -1)) for xt in xts]
        img = 4.0 * B * d / 1e6
        ym = 4.0 * B * op.ydim / 1e6
        outs = [torch.empty(B, d, device=dev) for _ in range(NBUF)]
        turn = [0]

        def rot(f):
            def call():
                k = turn[0] = (turn[0] + 1) % NBUF
                return f(k)
            return call

        calls = [
            ("project", rot(lambda k: op.project(xts[k], ys[k], out=outs[k])), 2 * img + ym),  # R x0, R y, W x0_hat
            ("ddnm_step", rot(lambda k: op.ddnm_step(xts[k], ets[k], zs[k], ys[k], 0.5, 0.6, 0.85, None)),
             5 * img + ym),                                                              # R xt et z y, W x0 x_next
            ("ddnm+_step", rot(lambda k: op.ddnm_step(xts[k], ets[k], zs[k], ys[k], 0.5, 0.6, 0.85, 0.1)), 5 * img + ym),
        ]
        for cname, fn, mb in calls:
            ms = timeit(fn)
            gbs = mb / ms
            print("%-22s %-12s %9.3f %9.1f %8.0f %6.2f" % (name, cname, ms, mb, gbs, gbs / peak))
    # evaluation metrics of a finished batch: two images in, per-image scalars out
    from nlc_b200 import metrics as M
    origs = [torch.rand(B, C, R, R, device=dev) for _ in range(NBUF)]
    samps = [(o + 0.05 * torch.randn_like(o)).clamp(0, 1) for o in origs]
    turn = [0]

    def nxt():
        turn[0] = (turn[0] + 1) % NBUF
        return turn[0]

    img = 4.0 * B * d / 1e6
    for cname, fn in (("ssim_fn", lambda k: M.ssim_fn(samps[k], origs[k])),
                      ("mse/psnr/l1", lambda k: M.restoration_metrics(xts[k], origs[k]))):
        ms = timeit(lambda: fn(nxt()))
        print("%-22s %-12s %9.3f %9.1f %8.0f %6.2f" % ("metrics", cname, ms, 2 * img, 2 * img / ms, 2 * img / ms / peak))

    # optimizer + EMA update of the sigma-model training step: 61.4 M parameters = the ADM-256 sigma-model
    import ctypes as Ct
    from nlc_b200 import _lib
    n = 61_400_000 // 4 * 4
    bufs = [torch.randn(n, device=dev) for _ in range(2)] + [torch.zeros(n, device=dev) for _ in range(2)] + \
        [torch.randn(n, device=dev)]
    L = _lib.lib()

    def adam():
        _lib.check(L.nlc_adamw_ema_step(_lib.ctx(0), bufs[0].data_ptr(), bufs[1].data_ptr(), bufs[2].data_ptr(),
                                        bufs[3].data_ptr(), bufs[4].data_ptr(), n, 1e-4, 0.9, 0.999, 1e-8, 0.01, 3, 0.999, 1.0,
                                        Ct.c_void_p(torch.cuda.current_stream().cuda_stream)))

    ms = timeit(adam)
    mb = 9 * 4.0 * n / 1e6  # p, g, m, v, ema read; p, m, v, ema written
    print("%-22s %-12s %9.3f %9.1f %8.0f %6.2f" % ("training", "adamw+ema", ms, mb, mb / ms, mb / ms / peak))

if __name__ == "__main__":
    main()
