"""Diagnose CUDA-graph replay time of one c2 timestep against eager launches, per batch size (device time by events)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import r02_report as RR  # noqa: E402

dev = torch.device("cuda:0")


def main():
    prec = sys.argv[1] if len(sys.argv) > 1 else "bf16"
    for B in (4, 32, 64, 256):
        exp, sch = RR.c2_experiment(prec)
        shape = (B, 3, 64, 64)
        w = exp._w(B)
        gb = w.graph_bufs()
        xT = torch.randn(shape, device=dev) * 50
        w.xa.copy_(xT)
        gb.step_sig.copy_(sch.sampling_sigmas[40:42])

        def body():
            eps, lv, s_t, s_p = exp.get_denoise_vector(w.xa, 0, gb.step_sig[0:1], gb.step_sig[1:2], "pred", True, True)
            x0 = exp._pred_xstart_clipped(w.xa, eps, s_t, w.x0)
            sch.pred_xprev(x0=x0, eps=eps, sigma_t=s_t, sigma_prev=s_p, xt=w.xa, log_variance=lv, noise=gb.noise,
                           out=w.xb, nan_flag=w.nan_flag)

        gb.noise.normal_()
        for _ in range(3):
            body()
        torch.cuda.synchronize()

        def timed(fn, n=20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            t1 = time.perf_counter()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n, (t1 - t0) * 1e3 / n

        eager_dev, eager_host = timed(body)
        t0 = time.perf_counter()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            body()
        torch.cuda.synchronize()
        t_cap = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter()
        g.replay()
        torch.cuda.synchronize()
        t_first = (time.perf_counter() - t0) * 1e3
        graph_dev, graph_host = timed(g.replay)
        print("B=%3d %s  eager: device %.3f ms/step (host issue %.3f)   graph: device %.3f ms/step (host %.3f)   capture %.1f ms, "
              "first replay %.1f ms" % (B, prec, eager_dev, eager_host, graph_dev, graph_host, t_cap, t_first), flush=True)
        del exp, g
        torch.cuda.empty_cache()


if __name__ == "__main__":
    with torch.no_grad():
        main()
