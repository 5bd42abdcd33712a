"""Round-2 accuracy report at the BENCHMARK architectures, against goldens recorded from the unmodified reference
(tests/golden/loop_c2_100.pt, nets_bench.pt, steps_adm256.pt; generator: tests/golden/make_golden.py bench_arch).

    python scripts/r02_report.py [c2] [nets] [adm]          (GPU box; never reads /root/reference)

c2  : the 100-step CelebA-64 NLC loop at batch 4, free-running, per operand mode: final-image PSNR (all samples and per
      sample), time-bucket flips (sample-steps whose t_hat = searchsorted(sigma_hat) differs from the reference's, counted up
      to each sample's first flip), sigma_hat error before the first flip, teacher-forced errors at the stored snapshots;
      CUDA-graph replay against the eager loop (identical result, time per loop).
nets: edm64 / adm256 network outputs per mode.    adm : c4 / c5 teacher-forced steps at 256 x 256.
"""
import math
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from nlc_b200 import synthetic_weights as weights  # noqa: E402

dev = torch.device("cuda:0")
GOLD = os.path.join(ROOT, "tests", "golden")
MODES = ("bf16", "fp16", "bf16+fp16", "tf32", "fp32")


def psnr(a, b):
    return 10 * math.log10(4.0 / max(torch.mean((a.double() - b.double()) ** 2).item(), 1e-30))


def l2rel(a, b):
    return (torch.linalg.vector_norm(a.double() - b.double()) / torch.linalg.vector_norm(b.double()).clamp_min(1e-30)).item()


def maxrel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


class Mixed:
    """forward in one precision, the sigma-determining encode pass in another (two model objects)."""

    def __init__(self, main, enc):
        self.main, self.enc = main, enc
        self.forward_scaled = main.forward_scaled
        self.encode_scaled = enc.encode_scaled


def c2_models(mode):
    from nlc_b200.unet_ddim import SigmaModel, UNetModel
    cfg = weights.CONFIGS["c2"]
    sd = weights.ddim_unet_state_dict(**cfg["unet"], seed=3)
    ssd = weights.ddim_sigma_state_dict(**cfg["sigma"], seed=4)
    pm, ps = (mode.split("+") + [None])[:2]
    m = UNetModel(**cfg["unet"], precision=pm, device=dev).load_state_dict(sd)
    if ps:
        m = Mixed(m, UNetModel(**cfg["unet"], precision=ps, device=dev).load_state_dict(sd))
    s = SigmaModel(**cfg["sigma"], precision=ps or pm, device=dev).load_state_dict(ssd)
    return m, s


def c2_experiment(mode):
    from nlc_b200.experiments import ImageExperiment
    from nlc_b200.schedulers import get_sampler
    m, s = c2_models(mode)
    sch = get_sampler("ddim_simple_orig", 1000, 100, start_sigma=100, eta=0.85).to(dev)
    exp = ImageExperiment(m, sch, batch_size=4, data_shape=(3, 64, 64), seed=5, device=dev)
    exp.set_model(m, s, learn_epsvar=False)
    exp.set_norm_maxmin(-2.0, 110.0)
    exp.set_clip_fn("clamp")
    return exp, sch


def c2_noise(shape, steps):
    torch.manual_seed(5)
    z = torch.randn(shape)
    return z, [torch.randn(shape) for _ in range(steps)]


def report_c2():
    gold = torch.load(os.path.join(GOLD, "loop_c2_100.pt"), weights_only=True)
    shape = (4, 3, 64, 64)
    z, noises = c2_noise(shape, 100)
    print("== c2 100-step loop, batch 4 (reference golden) ==")
    print("%-10s %8s  %-28s %6s %10s | teacher-forced worst: %9s %9s %9s" % (
        "mode", "PSNR", "per-sample PSNR", "flips", "sig err", "sigma_hat", "eps", "x_prev"))
    for mode in MODES:
        exp, sch = c2_experiment(mode)
        assert torch.equal(sch.timesteps.cpu(), gold["timesteps"]) and torch.equal(sch.sampling_sigmas.cpu(), gold["sigmas"])
        xT = (z / (1 / (gold["sigmas"][0] ** 2 + 1)).sqrt()).to(dev)
        sig_log = []

        def hook(ind, d):
            sig_log.append(d["sigma_t"].reshape(-1).expand(4).clone())

        out, _ = exp.denoise_loop(shape=shape, xT=xT, style="pred", norm_eps=True, refine_prior_sigma=True,
                                  return_log=False, sigma_pred_threshold=960, step_hook=hook,
                                  noise_fn=lambda i, like: noises[i].to(dev))
        sig = torch.stack(sig_log).cpu()  # [100, 4]
        t_hat = torch.searchsorted(gold["table"], sig.contiguous())
        diff = (t_hat != gold["t_hat"])
        first = [int(torch.nonzero(diff[:, b])[0]) if diff[:, b].any() else 100 for b in range(4)]
        n_flip = sum(1 for f in first if f < 100)
        serr = max(((sig[:max(f, 1), b] - gold["sigma_t"][:max(f, 1), b]).abs() / gold["sigma_t"][:max(f, 1), b]).max().item()
                   for b, f in enumerate(first))
        per = [psnr(out[b], gold["final"][b]) for b in range(4)]
        worst = dict(s=0.0, e=0.0, x=0.0)
        for i, sn in sorted(gold["snap"].items()):
            xt = sn["xt"].to(dev)
            t = int(gold["timesteps"][i])
            style, refine = ("pred", True) if t <= 960 else ("base", False)
            eps, lv, s_t, s_p = exp.get_denoise_vector(xt, t, sch.sampling_sigmas[i:i + 1], sch.sampling_sigmas[i + 1:i + 2],
                                                       style, True, refine)
            worst["s"] = max(worst["s"], l2rel(s_t.reshape(-1).cpu().expand(4), gold["sigma_t"][i]))
            worst["e"] = max(worst["e"], l2rel(eps.cpu(), sn["eps"]))
            x0h = sch.pred_xstart(xt, eps, s_t, clip=exp.clip_mode)
            xp = sch.pred_xprev(x0=x0h, eps=eps, sigma_t=s_t, sigma_prev=s_p, xt=xt, log_variance=lv, noise=noises[i].to(dev))
            worst["x"] = max(worst["x"], l2rel(xp.cpu(), sn["x_prev"]))
        print("%-10s %8.1f  %-28s %6d %10.2e | %30.2e %9.2e %9.2e   first flips at %s" % (
            mode, psnr(out, gold["final"]), " ".join("%.1f" % p for p in per), n_flip, serr, worst["s"], worst["e"],
            worst["x"], first))
        sys.stdout.flush()
        del exp
        torch.cuda.empty_cache()


def report_graph():
    from nlc_b200 import ops
    print("== CUDA-graph replay of the timestep vs the eager loop (c2, bf16) ==")
    for B in (4, 32, 256):
        exp, sch = c2_experiment("bf16")
        shape = (B, 3, 64, 64)
        g = torch.Generator().manual_seed(1)
        xT = (torch.randn(shape, generator=g) * (float(sch.sampling_sigmas[0]) ** 2 + 1) ** 0.5).to(dev)
        kw = dict(shape=shape, xT=xT, style="pred", norm_eps=True, refine_prior_sigma=True, return_log=False,
                  sigma_pred_threshold=960, to_cpu=False)
        res = {}
        for name, graph in (("eager", False), ("graph", True)):
            for rep in range(2):
                torch.cuda.manual_seed(7)
                ops.STATS.launches = 0
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                out, _ = exp.denoise_loop(graph=graph, **kw)
                torch.cuda.synchronize()
                dt = time.perf_counter() - t0
            res[name] = (out.clone(), dt, ops.STATS.launches)
        same = torch.equal(res["eager"][0], res["graph"][0])
        print("B=%3d  eager %.3f s (%d launches)   graph %.3f s (%d launches, %d replays)   identical=%s  max|diff|=%.2e" % (
            B, res["eager"][1], res["eager"][2], res["graph"][1], res["graph"][2], ops.STATS.graph_replays, same,
            (res["eager"][0] - res["graph"][0]).abs().max().item()))
        sys.stdout.flush()
        del exp
        torch.cuda.empty_cache()


ADM_KEYS = ("image_size", "model_channels", "out_channels", "num_res_blocks", "attention_resolutions", "channel_mult",
            "num_heads", "num_head_channels", "use_scale_shift_norm", "resblock_updown", "use_new_attention_order")


def adm_models(prec):
    from nlc_b200.unet_adm import SigmaModel, UNetModel
    cfg = dict(weights.ADM_CONFIGS["adm256"])
    sg = cfg.pop("sigma")
    m = UNetModel(in_channels=3, precision=prec, device=dev, **{k: cfg[k] for k in ADM_KEYS}).load_state_dict(
        weights.adm_unet_state_dict(**cfg, seed=3))
    s = SigmaModel(dim=sg["dim"], channels=sg["channels"], n_blocks=sg["n_blocks"], num_heads=cfg["num_heads"],
                   num_head_channels=cfg["num_head_channels"], use_new_attention_order=cfg["use_new_attention_order"],
                   precision=prec, device=dev).load_state_dict(weights.adm_sigma_state_dict(**sg, seed=4))
    return m, s


def report_nets():
    from nlc_b200.edm_networks import SigmaModel as ES, SongUNet
    gold = torch.load(os.path.join(GOLD, "nets_bench.pt"), weights_only=True)
    print("== networks at the benchmark architectures vs the reference (max-norm relative) ==")
    print("%-6s | edm64: %9s %9s %9s | adm256: %9s %9s %9s" % ("mode", "out", "feat", "r", "out", "feat", "r"))
    cfg = dict(weights.EDM_CONFIGS["edm64"])
    sg = cfg.pop("sigma")
    for prec in ("bf16", "fp16", "tf32", "fp32"):
        m = SongUNet(precision=prec, device=dev, **cfg).load_state_dict(weights.edm_unet_state_dict(**cfg, seed=3))
        s = ES(dim=sg["dim"], channels=sg["channels"], n_blocks=sg["n_blocks"], precision=prec, device=dev).load_state_dict(
            weights.edm_sigma_state_dict(**sg, seed=4))
        g = gold["edm64"]
        x, c = g["x"].to(dev), g["c_noise"].to(dev)
        e = [maxrel(m(x, c, None).cpu(), g["out"]), maxrel(m.encode(x, c, None).cpu(), g["feat"]),
             maxrel(s(g["feat"].to(dev)).cpu(), g["r"])]
        del m, s
        m, s = adm_models(prec)
        g = gold["adm256"]
        x, t = g["x"].to(dev), g["t"].to(dev)
        e += [maxrel(m(x, t).cpu(), g["out"]), maxrel(m.encode(x, t).cpu(), g["feat"]),
              maxrel(s(g["feat"].to(dev)).cpu(), g["r"])]
        print("%-6s | %16.2e %9.2e %9.2e | %17.2e %9.2e %9.2e" % ((prec,) + tuple(e)))
        sys.stdout.flush()
        del m, s
        torch.cuda.empty_cache()


def report_adm_steps():
    from functools import partial
    from nlc_b200 import constraint_functions as CF
    from nlc_b200.experiments import ImageExperiment
    from nlc_b200.schedulers import get_sampler
    gold = torch.load(os.path.join(GOLD, "steps_adm256.pt"), weights_only=True)
    print("== c4 / c5 teacher-forced steps at 256 x 256 (ADM-256, batch 1) ==")
    shape = (1, 3, 256, 256)
    torch.manual_seed(5)
    z = torch.randn(shape)
    noises = [torch.randn(shape) for _ in range(3)]
    for prec in ("bf16", "fp16", "tf32", "fp32"):
        m, s = adm_models(prec)
        for key, case in gold.items():
            task, scale = key.split("|")
            sch = get_sampler("ddim_simple_orig", 1000, 2, start_sigma=20.0, eta=0.85, sampler_var="learned").to(dev)
            exp = ImageExperiment(m, sch, batch_size=1, data_shape=shape[1:], seed=5, device=dev)
            exp.set_model(m, s, learn_epsvar=True)
            exp.set_norm_maxmin(-2.0, 110.0)
            exp.set_clip_fn("dynamic")
            con = CF.get_constraint_function(task, constraint_scale=float(scale), device=dev, image_size=256, channels=3)
            y = case["y"].to(dev)
            cfn = partial(con.constraint_fn, y=y, lambda_t=con.lr)
            w = exp._w(1)
            xt = (z / (1 / (case["sigmas"][0] ** 2 + 1)).sqrt()).to(dev)
            errs = []
            for i in range(len(case["x_prev"])):
                t = int(case["timesteps"][i])
                eps, lv, s_t, s_p = exp.get_denoise_vector(xt, t, sch.sampling_sigmas[i:i + 1], sch.sampling_sigmas[i + 1:i + 2],
                                                           "pred", True, True)
                x0h = exp._pred_xstart_clipped(xt, eps, s_t, w.x0)
                x0 = cfn(x0h)
                xp = sch.pred_xprev(x0=x0, eps=eps, sigma_t=s_t, sigma_prev=s_p, xt=xt, log_variance=lv,
                                    noise=noises[i].to(dev))
                errs.append((l2rel(s_t.reshape(-1).cpu(), case["sigma_t"][i]), l2rel(x0.cpu(), case["x0"][i]),
                             l2rel(xp.cpu(), case["x_prev"][i])))
                xt = case["x_prev"][i].to(dev)  # teacher forcing: the reference's own x_{t-1}
            print("%-5s %-22s " % (prec, key) + "  ".join("step%d: sig %.1e x0 %.1e xprev %.1e" % ((i,) + e)
                                                           for i, e in enumerate(errs)))
            sys.stdout.flush()
        del m, s
        torch.cuda.empty_cache()


if __name__ == "__main__":
    what = sys.argv[1:] or ["c2", "graph", "nets", "adm"]
    with torch.no_grad():
        if "c2" in what:
            report_c2()
        if "graph" in what:
            report_graph()
        if "nets" in what:
            report_nets()
        if "adm" in what:
            report_adm_steps()
