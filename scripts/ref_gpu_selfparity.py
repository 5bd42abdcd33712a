"""How closely does the REFERENCE's own GPU path reproduce the reference's CPU run?  (test-infrastructure script: runs
the oracle, which is pinned torch.equal to the reference, through PyTorch eager / cuDNN on the GPU.)

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

from oracle import ddim_net, sampler as S, weights  # noqa: E402


def psnr(a, b):
    return 10 * math.log10(4.0 / max(torch.mean((a.double().cpu() - b.double().cpu()) ** 2).item(), 1e-30))


def main():
    dev = torch.device("cuda:0")
    g = torch.load(os.path.join(ROOT, "tests", "golden", "loop_c2_100.pt"), weights_only=True)
    shape = (4, 3, 64, 64)
    torch.manual_seed(5)
    z = torch.randn(shape)
    noises = [torch.randn(shape) for _ in range(100)]
    cfg = weights.CONFIGS["c2"]
    sd = {k: v.to(dev) for k, v in weights.ddim_unet_state_dict(**cfg["unet"], seed=3).items()}
    ssd = {k: v.to(dev) for k, v in weights.ddim_sigma_state_dict(**cfg["sigma"], seed=4).items()}
    tab = S.Tables()
    ts, sig, mvc = tab.ddim_schedule(100.0, None, 100)
    assert torch.equal(ts, g["timesteps"]) and torch.equal(sig, g["sigmas"])
    for name in ("betas", "alphas_cumprod", "sigmas", "posterior_variance"):
        setattr(tab, name, getattr(tab, name).to(dev))
    mvc = mvc.to(dev) if torch.is_tensor(mvc) else mvc
    sig_d = sig.to(dev)
    d = 3 * 64 * 64
    xT = (z / (1 / (g["sigmas"][0] ** 2 + 1)).sqrt()).to(dev)
    noises_d = [n.to(dev) for n in noises]
    torch.set_default_device(dev)

    fwd = lambda z_, t: ddim_net.unet_forward(sd, z_, t)
    enc = lambda z_, t: ddim_net.unet_encode(sd, z_, t)
    sgf = lambda f: ddim_net.sigma_forward(ssd, f)

    class ForcedTables:
        """The golden run's time buckets: the two lookups of step i return g['t_first'][i] / g['t_hat'][i]."""

        def __init__(self, inner):
            self.inner, self.calls = inner, 0

        def __getattr__(self, k):
            return getattr(self.inner, k)

        def t_of_sigma(self, sigma):
            i, which = divmod(self.calls, 2)
            self.calls += 1
            return (g["t_first"][i] if which == 0 else g["t_hat"][i]).to(dev)

    settings = [("fp32 cuDNN (allow_tf32 off)", False, None),
                ("TF32 convolutions (the reference's default GPU run)", True, None),
                ("torch.autocast(float16)", True, torch.float16)]
    for label, tf32, autocast in settings:
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        res = {}
        for forced in (True, False):
            log = []
            t_tab = ForcedTables(tab) if forced else tab
            with torch.no_grad(), torch.autocast("cuda", dtype=autocast, enabled=autocast is not None):
                # (steps above sigma_pred_threshold use style 'base': one lookup only - none in this schedule's range
                # matters for the forced table because refine is off there and t is the schedule's own)
                out = S.denoise_loop(t_tab, ts.tolist(), sig_d, mvc, fwd, enc, sgf, xT, kind="ddim_simple_orig", eta=0.85,
                                     style="pred", norm_eps=True, refine=True, norm_min=-2.0 / d ** 0.5,
                                     norm_max=110.0 / d ** 0.5, noises=noises_d, sigma_pred_threshold=960, log=log)
            sig_log = torch.stack([s["sigma_t"].reshape(-1).expand(4) for s in log]).float().cpu()
            flips = int((torch.searchsorted(g["table"], sig_log.contiguous()) != g["t_hat"]).any(dim=0).sum())
            res[forced] = (psnr(out.float(), g["final"]), flips)
        print("%-55s PSNR %.1f dB with the CPU run's time buckets, %.1f dB entirely free (%d of 4 samples crossed a time "
              "bucket)" % (label, res[True][0], res[False][0], res[False][1]), flush=True)


if __name__ == "__main__":
    main()
