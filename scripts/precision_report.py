"""Accuracy of the four operand modes against the reference's golden dumps, in one table.

    python scripts/precision_report.py            (GPU box; reads tests/golden/*.pt, never /root/reference)

Per mode (bf16, fp16, tf32, fp32): max-norm relative error of the UNet output (tiny golden case and the c2 architecture
against the CPU oracle), worst teacher-forced per-step L2-relative error of sigma_hat / eps / x_{t-1} over the eight
schedulers of tests/golden/denoise_loop_tiny.pt, and the final-image PSNR (peak-to-peak 2) of the free-running
ddim_simple_orig loop and of the four DDNM-constrained ADM loops (tests/golden/loops3_constrained.pt).
"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle import ddim_net, weights  # noqa: E402
import test_gpu_constrained as TC  # noqa: E402
import test_gpu_sampler as TS  # noqa: E402

dev = torch.device("cuda:0")
GOLD = os.path.join(ROOT, "tests", "golden")


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def psnr(a, b):
    return 10 * math.log10(4.0 / max(torch.mean((a.double() - b.double()) ** 2).item(), 1e-30))


def main():
    from nlc_b200.unet_ddim import UNetModel
    nets = torch.load(os.path.join(GOLD, "nets_tiny.pt"), weights_only=True)
    loops = torch.load(os.path.join(GOLD, "denoise_loop_tiny.pt"), weights_only=True)
    cons = torch.load(os.path.join(GOLD, "loops3_constrained.pt"), weights_only=True)
    cfg2 = weights.CONFIGS["c2"]
    sd2 = weights.ddim_unet_state_dict(**cfg2["unet"], seed=3)
    g = torch.Generator().manual_seed(5)
    x2 = torch.randn(2, 3, 64, 64, generator=g)
    t2 = torch.tensor([999.0, 250.0])
    with torch.no_grad():
        ref2 = ddim_net.unet_forward(sd2, x2, t2)
    print("%-5s %10s %10s | %10s %10s %10s | %9s %s" % ("mode", "net tiny", "net c2", "sigma_hat", "eps", "x_prev",
                                                        "PSNR loop", "PSNR constrained (sr, inpaint, color, cs)"))
    for prec in ("bf16", "fp16", "tf32", "fp32"):
        cfg = weights.CONFIGS["tiny"]
        m = UNetModel(**cfg["unet"], precision=prec, device=dev).load_state_dict(
            weights.ddim_unet_state_dict(**cfg["unet"], seed=3))
        e_tiny = rel(m(nets["x"].to(dev), nets["t"].to(dev)).cpu(), nets["out"])
        m2 = UNetModel(**cfg2["unet"], precision=prec, device=dev).load_state_dict(sd2)
        e_c2 = rel(m2(x2.to(dev), t2.to(dev)).cpu(), ref2)
        del m2
        worst = dict(sigma=0.0, eps=0.0, x_prev=0.0)
        for key, case in loops.items():
            kind, eta, var = key.split("|")
            exp, sch = TS._setup(prec, kind, float(eta), var)
            for i in range(len(case["eps"])):
                xt = case["xt"][i].to(dev)
                eps, lv, s_t, s_p = exp.get_denoise_vector(xt, int(case["timesteps"][i]), sch.sampling_sigmas[i:i + 1],
                                                           sch.sampling_sigmas[i + 1:i + 2], "pred", True, True)
                worst["sigma"] = max(worst["sigma"], TS._l2rel(s_t.reshape(-1).cpu(), case["sigma_t"][i]))
                worst["eps"] = max(worst["eps"], TS._l2rel(eps.cpu(), case["eps"][i]))
                x0h = sch.pred_xstart(xt, eps, s_t, clip=exp.clip_mode)
                noise = case["noises"][i].to(dev) if case["noises"] else None
                xp = sch.pred_xprev(x0=x0h, eps=eps, sigma_t=s_t, sigma_prev=s_p, xt=xt, log_variance=lv, noise=noise)
                worst["x_prev"] = max(worst["x_prev"], TS._l2rel(xp.cpu(), case["x_prev"][i]))
        case = loops["ddim_simple_orig|0.85|none"]
        exp, sch = TS._setup(prec, "ddim_simple_orig", 0.85, "none")
        xT = (case["z"] / (1 / (case["sigmas"][0] ** 2 + 1)).sqrt()).to(dev)
        out, _ = exp.denoise_loop(shape=tuple(xT.shape), xT=xT, style="pred", norm_eps=True, refine_prior_sigma=True,
                                  return_log=False, noise_fn=lambda i, like: case["noises"][i].to(dev))
        p_loop = psnr(out, case["final"])
        p_con = []
        for key in TC.TASKS:
            exp, sch, con, c, y, cfn, closs = TC._setup(prec, cons, key)
            xT = (c["z"] / (1 / (c["sigmas"][0] ** 2 + 1)).sqrt()).to(dev)
            out, _ = exp.denoise_loop(shape=tuple(xT.shape), xT=xT, style="pred", constrain_fn=cfn, norm_eps=True,
                                      refine_prior_sigma=True, return_log=False, chunk_size=1, constrain_loss=closs,
                                      sigma_pred_threshold=960, noise_fn=lambda i, like: c["noises"][i].to(dev))
            p_con.append(psnr(out, c["final"]))
        print("%-5s %10.2e %10.2e | %10.2e %10.2e %10.2e | %9.1f %s" % (
            prec, e_tiny, e_c2, worst["sigma"], worst["eps"], worst["x_prev"], p_loop,
            " ".join("%.1f" % v for v in p_con)))


if __name__ == "__main__":
    main()
