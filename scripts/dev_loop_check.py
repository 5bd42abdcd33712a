"""Development check on a B200: the GPU denoise_loop against the CPU oracle loop (free-running and teacher-forced)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nlc_b200.experiments import ImageExperiment
from nlc_b200.schedulers import get_sampler
from nlc_b200.unet_ddim import SigmaModel, UNetModel
from oracle import ddim_net, sampler as S, weights

dev = torch.device("cuda:0")


def relerr(a, b):
    return (torch.linalg.vector_norm(a - b) / torch.linalg.vector_norm(b).clamp_min(1e-30)).item()


def run(name, precision, kind, eta, var, n_steps=8, B=2, start_sigma=40.0):
    cfg = weights.CONFIGS[name]
    R = cfg["unet"]["image_size"]
    sd = weights.ddim_unet_state_dict(**cfg["unet"], seed=3)
    ssd = weights.ddim_sigma_state_dict(**cfg["sigma"], seed=4)
    shape = (B, 3, R, R)
    d = 3 * R * R
    tab = S.Tables()
    ts, sig, mvc = tab.ddim_schedule(start_sigma, None, n_steps)
    g = torch.Generator().manual_seed(7)
    xT = torch.randn(shape, generator=g) / (1 / (sig[0] ** 2 + 1)).sqrt()
    noises = [torch.randn(shape, generator=g) for _ in range(len(ts) - 1)]
    nmin, nmax = -2.0, 0.9 * d ** 0.5
    log = []
    fwd = lambda z, t: ddim_net.unet_forward(sd, z, t)
    enc = lambda z, t: ddim_net.unet_encode(sd, z, t)
    sgf = lambda f: ddim_net.sigma_forward(ssd, f)
    with torch.no_grad():
        ref = S.denoise_loop(tab, ts.tolist(), sig, mvc, fwd, enc, sgf, xT, kind=kind, eta=eta, sampler_var=var,
                             style="pred", norm_eps=True, refine=True, norm_min=nmin / d ** 0.5, norm_max=nmax / d ** 0.5,
                             noises=noises, log=log)
    model = UNetModel(**cfg["unet"], precision=precision, device=dev).load_state_dict(sd)
    smodel = SigmaModel(**cfg["sigma"], precision=precision, device=dev).load_state_dict(ssd)
    sch = get_sampler(kind, 1000, n_steps, start_sigma=start_sigma, eta=eta, sampler_var=var).to(dev)
    assert torch.equal(sch.timesteps.cpu(), ts) and torch.equal(sch.sampling_sigmas.cpu(), sig)
    exp = ImageExperiment(model, sch, batch_size=B, data_shape=(3, R, R), seed=0, device=dev)
    exp.set_model(model, smodel, learn_epsvar=False)
    exp.set_norm_maxmin(nmin, nmax)
    exp.set_clip_fn("clamp")
    got = []
    out, _ = exp.denoise_loop(shape=shape, xT=xT.to(dev), style="pred", norm_eps=True, refine_prior_sigma=True,
                              return_log=False, noise_fn=lambda i, like: noises[i].to(dev),
                              step_hook=lambda i, dct: got.append({k: v.detach().clone().cpu() for k, v in dct.items()}))
    free = relerr(out, ref)
    # teacher-forced: feed the oracle's x_t of every step through one GPU step
    worst = {"eps": 0.0, "x_prev": 0.0, "sigma_t": 0.0, "x0_hat": 0.0}
    for i, st in enumerate(log):
        sch.reset_state()
        xt = st["xt"].to(dev)
        eps, lv, s_t, s_p = exp.get_denoise_vector(xt, int(ts[i]), sch.sampling_sigmas[i:i + 1],
                                                   sch.sampling_sigmas[i + 1:i + 2], "pred", True, True)
        x0h = sch.pred_xstart(xt, eps, s_t, clip=exp.clip_mode)
        xp = sch.pred_xprev(x0=x0h, eps=eps, sigma_t=s_t, sigma_prev=s_p, xt=xt, log_variance=lv,
                            noise=noises[i].to(dev))
        worst["eps"] = max(worst["eps"], relerr(eps.cpu(), st["eps"]))
        worst["x0_hat"] = max(worst["x0_hat"], relerr(x0h.cpu(), st["x0_hat"]))
        worst["x_prev"] = max(worst["x_prev"], relerr(xp.cpu(), st["x_prev"]))
        worst["sigma_t"] = max(worst["sigma_t"], relerr(s_t.reshape(-1).cpu(), st["sigma_t"].reshape(-1)))
    mse = torch.mean((out - ref) ** 2).item()
    psnr = 10 * torch.log10(torch.tensor(4.0 / max(mse, 1e-20))).item()
    print("%-5s %-4s %-16s eta %.2f %-10s free-run rel %.2e PSNR %.1f dB | teacher-forced rel: eps %.2e x0 %.2e "
          "x_prev %.2e sigma %.2e" % (name, precision, kind, eta, var, free, psnr, worst["eps"], worst["x0_hat"],
                                      worst["x_prev"], worst["sigma_t"]), flush=True)


if __name__ == "__main__":
    for prec in ("tf32", "bf16"):
        run("tiny", prec, "ddim", 0.0, "none")
        run("tiny", prec, "ddim_simple_orig", 0.85, "none")
        run("tiny", prec, "ddim", 0.5, "fixedsmall")
        run("tiny", prec, "ddpm", 1.0, "fixedlarge")
        run("tiny", prec, "ddpm_orig", 1.0, "fixedsmall")
        run("tiny", prec, "ddim_orig", 0.3, "fixedlarge")
        run("tiny", prec, "ddim_simple", 0.2, "none")
        run("tiny", prec, "ddim_simple_drag", 0.2, "none")
        run("c1", prec, "ddim_simple_orig", 0.85, "none", n_steps=20, B=4, start_sigma=100.0)
        run("c2", prec, "ddim_simple_orig", 0.85, "none", n_steps=10, B=2, start_sigma=100.0)
