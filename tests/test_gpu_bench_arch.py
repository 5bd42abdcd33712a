"""GPU parity at the BENCHMARK architectures against the unmodified reference (tests/golden/loop_c2_100.pt, nets_bench.pt,
steps_adm256.pt; generator: tests/golden/make_golden.py bench_arch): config c2 (CelebA-64 unet_ddim, the 100-step NLC
loop at batch 4), c3 (EDM SongUNet-64 network), c4 / c5 (ADM-256 network; DDNM-constrained steps at 256 x 256).

Gates (L2-relative per step unless stated; measured values: profiles/r02_report_*.log):
  fp32 mode (3 x tf32 split products): every per-step tensor <= 1e-4 (the north-star's fp32/tf32 tolerance), free-running
      100-step final image >= 100 dB, no time-bucket flip.
  tf32 / fp16 modes: teacher-forced sigma_hat 1e-3, eps / x_{t-1} 5e-3.
  bf16 mode:         teacher-forced sigma_hat 8e-3, eps 6e-2, x_{t-1} 1.5e-2 (measured 1.05e-2 on the ADM-256 SR step).
  Free-running 100-step loop, 16-bit and tf32 modes: the final-image PSNR against the reference's (fp32, CPU) run has two
      gates.  (a) With the reference's own time buckets (both discrete lookups t = searchsorted(sigma) of every step taken
      from the recorded reference run, everything else free-running - `ExperimentDiffusion.time_source`): >= 45 dB in every
      mode, the north-star's 16-bit gate; this is the number arithmetic precision decides (measured: fp16 63, tf32 66, bf16
      46 dB).  (b) Entirely free: in EVERY reduced-precision arithmetic all four samples cross a ~1 %-wide time-bucket edge
      somewhere in the 400 sample-steps, after which a random-init network moves that sample's eps by ~1e-2 and the
      trajectories separate chaotically - including the reference's OWN default GPU run (TF32 convolutions through cuDNN),
      which ends 44-46 dB from its CPU run (torch.autocast fp16: 43.7 dB; profiles/r02h_ref_gpu_selfparity.log).  "Matches
      the reference" cannot be gated tighter than the reference matches itself, so the free gate is relative to that
      CONTROL, run here through the oracle on the GPU (parity_util.oracle_c2_loop_on_gpu): fp16 / tf32 within 6 dB of the
      control and >= 38 dB, bf16 >= 33 dB; the flip counts are printed."""
import math
import os
from functools import partial

import pytest
import torch

from oracle import weights
from test_oracle_golden import bench_noise, same_digest

pytestmark = pytest.mark.gpu
dev = torch.device("cuda:0")
ADM_KEYS = ("image_size", "model_channels", "out_channels", "num_res_blocks", "attention_resolutions", "channel_mult",
            "num_heads", "num_head_channels", "use_scale_shift_norm", "resblock_updown", "use_new_attention_order")
STEP_TOL = {"fp32": dict(sigma=1e-4, eps=1e-4, x=1e-4), "tf32": dict(sigma=1e-3, eps=5e-3, x=5e-3),
            "fp16": dict(sigma=1e-3, eps=5e-3, x=5e-3), "bf16": dict(sigma=8e-3, eps=6e-2, x=1.5e-2)}
NET_TOL = {"fp32": 1e-4, "tf32": 5e-3, "fp16": 5e-3, "bf16": 3e-2}  # max-norm relative, network outputs
PSNR_SYNC = {"fp32": 100.0, "tf32": 45.0, "fp16": 45.0, "bf16": 45.0}
PSNR_FREE = {"fp32": 100.0, "tf32": 38.0, "fp16": 38.0, "bf16": 33.0}  # absolute floors; fp16 / tf32 also: control - 6 dB
CONTROL_MARGIN = 6.0


def _l2rel(a, b):
    return (torch.linalg.vector_norm(a.double() - b.double()) / torch.linalg.vector_norm(b.double()).clamp_min(1e-30)).item()


def _maxrel(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


def _psnr(a, b):
    return 10 * math.log10(4.0 / max(torch.mean((a.double() - b.double()) ** 2).item(), 1e-30))


@pytest.fixture(scope="module")
def c2_gold(golden_dir):
    g = torch.load(os.path.join(golden_dir, "loop_c2_100.pt"), weights_only=True)
    z, noises = bench_noise((4, 3, 64, 64), 100)
    assert same_digest(z, g["z_digest"])
    assert all(same_digest(n, d) for n, d in zip(noises, g["noise_digest"]))
    return g, z, noises


@pytest.fixture(scope="module")
def ref_gpu_control(c2_gold):
    """Free-running PSNR of the reference's own default GPU arithmetic (oracle, cuDNN TF32 convolutions) against its CPU run."""
    from parity_util import oracle_c2_loop_on_gpu
    g, z, noises = c2_gold
    out, flips = oracle_c2_loop_on_gpu(g, z, noises, tf32=True, forced=False)
    p = _psnr(out, g["final"])
    print("\ncontrol - the reference's default GPU run (TF32 cuDNN) against its CPU run: %.1f dB entirely free (%d of 4 "
          "samples crossed a time bucket)" % (p, flips))
    return p


def _c2_experiment(prec):
    from nlc_b200.experiments import ImageExperiment
    from nlc_b200.schedulers import get_sampler
    from nlc_b200.unet_ddim import SigmaModel, UNetModel
    cfg = weights.CONFIGS["c2"]
    m = UNetModel(**cfg["unet"], precision=prec, device=dev).load_state_dict(
        weights.ddim_unet_state_dict(**cfg["unet"], seed=3))
    s = SigmaModel(**cfg["sigma"], precision=prec, device=dev).load_state_dict(
        weights.ddim_sigma_state_dict(**cfg["sigma"], seed=4))
    sch = get_sampler("ddim_simple_orig", 1000, 100, start_sigma=100, eta=0.85).to(dev)
    exp = ImageExperiment(m, sch, batch_size=4, data_shape=(3, 64, 64), seed=5, device=dev)
    exp.set_model(m, s, learn_epsvar=False)
    exp.set_norm_maxmin(-2.0, 110.0)
    exp.set_clip_fn("clamp")
    return exp, sch


@pytest.mark.parametrize("prec", ["fp32", "tf32", "fp16", "bf16"])
def test_c2_teacher_forced_snapshots(c2_gold, prec):
    """The reference's x_t of eight steps of its own 100-step c2 run through one GPU step each."""
    g, z, noises = c2_gold
    exp, sch = _c2_experiment(prec)
    assert torch.equal(sch.timesteps.cpu(), g["timesteps"]) and torch.equal(sch.sampling_sigmas.cpu(), g["sigmas"])
    tol = STEP_TOL[prec]
    for i, sn in sorted(g["snap"].items()):
        xt = sn["xt"].to(dev)
        t = int(g["timesteps"][i])
        style, refine = ("pred", True) if t <= 960 else ("base", False)
        eps, lv, s_t, s_p = exp.get_denoise_vector(xt, t, sch.sampling_sigmas[i:i + 1], sch.sampling_sigmas[i + 1:i + 2],
                                                   style, True, refine)
        s_ours = s_t.reshape(-1).cpu().expand(4)
        assert _l2rel(s_ours, g["sigma_t"][i]) < tol["sigma"], (prec, i)
        same = torch.searchsorted(g["table"], s_ours.contiguous()) == g["t_hat"][i]
        if prec == "fp32":
            assert same.all(), (i, "time-bucket flip in the fp32 mode")
        x0h = sch.pred_xstart(xt, eps, s_t, clip=exp.clip_mode)
        xp = sch.pred_xprev(x0=x0h, eps=eps, sigma_t=s_t, sigma_prev=s_p, xt=xt, log_variance=lv, noise=noises[i].to(dev))
        for b in range(4):  # per sample: the stated tolerance when its time bucket is the reference's, 1e-1 when it is not
            lim_e, lim_x = (tol["eps"], tol["x"]) if same[b] else (1e-1, 1e-1)
            assert _l2rel(eps[b].cpu(), sn["eps"][b]) < lim_e, (prec, i, b)
            assert _l2rel(xp[b].cpu(), sn["x_prev"][b]) < lim_x, (prec, i, b)


def _run_c2(exp, sch, g, z, noises, forced):
    xT = (z / (1 / (g["sigmas"][0] ** 2 + 1)).sqrt()).to(dev)
    sig_log = []
    exp.time_source = (lambda i: (g["t_first"][i].to(dev), g["t_hat"][i].float().to(dev))) if forced else None
    out, _ = exp.denoise_loop(shape=(4, 3, 64, 64), xT=xT, style="pred", norm_eps=True, refine_prior_sigma=True,
                              return_log=False, sigma_pred_threshold=960,
                              step_hook=lambda i, d: sig_log.append(d["sigma_t"].reshape(-1).expand(4).clone()),
                              noise_fn=lambda i, like: noises[i].to(dev))
    exp.time_source = None
    sig = torch.stack(sig_log).cpu()
    flips = int((torch.searchsorted(g["table"], sig.contiguous()) != g["t_hat"]).any(dim=0).sum())
    return out, flips


@pytest.mark.parametrize("prec", ["fp32", "tf32", "fp16", "bf16"])
def test_c2_free_running_100_steps(c2_gold, ref_gpu_control, prec):
    g, z, noises = c2_gold
    exp, sch = _c2_experiment(prec)
    out, _ = _run_c2(exp, sch, g, z, noises, forced=True)
    p_sync = _psnr(out, g["final"])
    out, flips = _run_c2(exp, sch, g, z, noises, forced=False)
    p_free = _psnr(out, g["final"])
    print("\nc2 100 steps %s: PSNR %.1f dB with the reference's time buckets, %.1f dB entirely free (%d of 4 samples "
          "crossed a time bucket)" % (prec, p_sync, p_free, flips))
    assert p_sync >= PSNR_SYNC[prec], (prec, p_sync)
    assert p_free >= PSNR_FREE[prec], (prec, p_free, flips)
    if prec in ("tf32", "fp16"):
        assert p_free >= ref_gpu_control - CONTROL_MARGIN, (prec, p_free, ref_gpu_control, flips)
    if prec == "fp32":
        assert flips == 0


def test_c2_graph_replay_equals_eager_loop(c2_gold):
    """The CUDA-graph replay of the timestep is the eager loop, bit for bit (same kernels, same order of random draws)."""
    g, z, noises = c2_gold
    exp, sch = _c2_experiment("fp16")
    xT = (z / (1 / (g["sigmas"][0] ** 2 + 1)).sqrt()).to(dev)
    outs = []
    for graph in (False, True):
        torch.cuda.manual_seed(11)
        out, _ = exp.denoise_loop(shape=(4, 3, 64, 64), xT=xT, style="pred", norm_eps=True, refine_prior_sigma=True,
                                  return_log=False, sigma_pred_threshold=960, graph=graph, to_cpu=False)
        outs.append(out.clone())
    assert torch.equal(outs[0], outs[1])
    # the captured timestep is kept on the experiment: a second call (other start, other noise) replays it from its first
    # capturable step and still equals the eager loop; a changed host scalar (eta) is a different signature
    assert len(exp._graph_cache) == 1
    xT2 = xT.flip(0) * 0.9
    for eta in (0.85, 0.5):
        sch.eta = eta
        outs = []
        for graph in (False, True):
            torch.cuda.manual_seed(12)
            out, _ = exp.denoise_loop(shape=(4, 3, 64, 64), xT=xT2, style="pred", norm_eps=True, refine_prior_sigma=True,
                                      return_log=False, sigma_pred_threshold=960, graph=graph, to_cpu=False)
            outs.append(out.clone())
        assert torch.equal(outs[0], outs[1]), eta
    assert len(exp._graph_cache) == 2
    exp.clear_graphs()


# ------------------------------------------------------------------------------------------------------ networks
@pytest.fixture(scope="module")
def nets_gold(golden_dir):
    return torch.load(os.path.join(golden_dir, "nets_bench.pt"), weights_only=True)


@pytest.mark.parametrize("prec", ["fp32", "tf32", "fp16", "bf16"])
def test_edm64_network(nets_gold, prec):
    from nlc_b200.edm_networks import SigmaModel, SongUNet
    cfg = dict(weights.EDM_CONFIGS["edm64"])
    sg = cfg.pop("sigma")
    m = SongUNet(precision=prec, device=dev, **cfg).load_state_dict(weights.edm_unet_state_dict(**cfg, seed=3))
    s = SigmaModel(dim=sg["dim"], channels=sg["channels"], n_blocks=sg["n_blocks"], precision=prec,
                   device=dev).load_state_dict(weights.edm_sigma_state_dict(**sg, seed=4))
    g = nets_gold["edm64"]
    x, c = g["x"].to(dev), g["c_noise"].to(dev)
    tol = NET_TOL[prec]
    assert _maxrel(m(x, c, None).cpu(), g["out"]) < tol
    assert _maxrel(m.encode(x, c, None).cpu(), g["feat"]) < tol
    # r is a correction of order 1e-1 to 1: judged on the scale of 1 + r (sigma_hat = sigma (1 + r))
    assert (s(g["feat"].to(dev)).cpu() - g["r"]).abs().max() < tol


@pytest.mark.parametrize("prec", ["fp32", "tf32", "fp16", "bf16"])
@pytest.mark.parametrize("name", ["dhariwal_tiny", "dhariwal64"])
def test_dhariwal_unet(golden_dir, prec, name):
    """DhariwalUNet (src/edm_networks.py:406-502: adaptive scale/shift, 64-channel heads at 32 / 16 / 8, weight-less
    resampling skips) against the unmodified reference, at a two-level shape and at 64 x 64 with four levels."""
    from nlc_b200.edm_networks import DhariwalUNet
    cfg = dict(weights.DHARIWAL_CONFIGS[name])
    g = torch.load(os.path.join(golden_dir, "nets_dhariwal.pt"), weights_only=True)[name]
    m = DhariwalUNet(precision=prec, device=dev, **cfg).load_state_dict(weights.dhariwal_unet_state_dict(**cfg, seed=3))
    out = m(g["x"].to(dev), g["c_noise"].to(dev), None)
    assert _maxrel(out.cpu(), g["out"]) < NET_TOL[prec]
    with pytest.raises(AttributeError):
        m.encode(g["x"].to(dev), g["c_noise"].to(dev), None)


def _adm_models(prec):
    from nlc_b200.unet_adm import SigmaModel, UNetModel
    cfg = dict(weights.ADM_CONFIGS["adm256"])
    sg = cfg.pop("sigma")
    m = UNetModel(in_channels=3, precision=prec, device=dev, **{k: cfg[k] for k in ADM_KEYS}).load_state_dict(
        weights.adm_unet_state_dict(**cfg, seed=3))
    s = SigmaModel(dim=sg["dim"], channels=sg["channels"], n_blocks=sg["n_blocks"], num_heads=cfg["num_heads"],
                   num_head_channels=cfg["num_head_channels"], use_new_attention_order=cfg["use_new_attention_order"],
                   precision=prec, device=dev).load_state_dict(weights.adm_sigma_state_dict(**sg, seed=4))
    return m, s


@pytest.fixture(scope="module")
def adm_gold(golden_dir):
    return torch.load(os.path.join(golden_dir, "steps_adm256.pt"), weights_only=True)


@pytest.mark.parametrize("prec", ["fp32", "tf32", "fp16", "bf16"])
def test_adm256_network_and_constrained_steps(nets_gold, adm_gold, prec):
    """ADM-256 (552.8 M parameters) forward / encode / sigma-model against the reference, then the c4 (SR x4) and c5
    (colourisation) DDNM-constrained NLC steps at 256 x 256, teacher-forced on the reference's own x_t."""
    from nlc_b200 import constraint_functions as CF
    from nlc_b200.experiments import ImageExperiment
    from nlc_b200.schedulers import get_sampler
    m, s = _adm_models(prec)
    g = nets_gold["adm256"]
    x, t = g["x"].to(dev), g["t"].to(dev)
    tol = NET_TOL[prec]
    assert _maxrel(m(x, t).cpu(), g["out"]) < tol
    assert _maxrel(m.encode(x, t).cpu(), g["feat"]) < tol
    assert (s(g["feat"].to(dev)).cpu() - g["r"]).abs().max() < tol
    shape = (1, 3, 256, 256)
    z, noises = bench_noise(shape, 3)
    st = STEP_TOL[prec]
    for key, case in adm_gold.items():
        assert same_digest(z, case["z_digest"])
        assert all(same_digest(n, d) for n, d in zip(noises, case["noise_digest"]))
        task, scale = key.split("|")
        sch = get_sampler("ddim_simple_orig", 1000, 2, start_sigma=20.0, eta=0.85, sampler_var="learned").to(dev)
        assert torch.equal(sch.timesteps.cpu(), case["timesteps"])
        exp = ImageExperiment(m, sch, batch_size=1, data_shape=shape[1:], seed=5, device=dev)
        exp.set_model(m, s, learn_epsvar=True)
        exp.set_norm_maxmin(-2.0, 110.0)
        exp.set_clip_fn("dynamic")
        con = CF.get_constraint_function(task, constraint_scale=float(scale), device=dev, image_size=256, channels=3)
        y = case["y"].to(dev)
        x_true = torch.rand(shape, generator=torch.Generator().manual_seed(32)) * 2 - 1
        assert same_digest(x_true, case["x_true_digest"])
        assert (con.transform(x_true.to(dev)).cpu().reshape(1, -1) - case["y"].reshape(1, -1)).abs().max() < 1e-5
        cfn = partial(con.constraint_fn, y=y, lambda_t=con.lr)
        closs = partial(con.loss, y=y)
        w = exp._w(1)
        xt = (z / (1 / (case["sigmas"][0] ** 2 + 1)).sqrt()).to(dev)
        for i in range(len(case["x_prev"])):
            eps, lv, s_t, s_p = exp.get_denoise_vector(xt, int(case["timesteps"][i]), sch.sampling_sigmas[i:i + 1],
                                                       sch.sampling_sigmas[i + 1:i + 2], "pred", True, True)
            assert _l2rel(s_t.reshape(-1).cpu(), case["sigma_t"][i]) < st["sigma"], (prec, key, i)
            x0 = cfn(exp._pred_xstart_clipped(xt, eps, s_t, w.x0))
            xp = sch.pred_xprev(x0=x0, eps=eps, sigma_t=s_t, sigma_prev=s_p, xt=xt, log_variance=lv,
                                noise=noises[i].to(dev))
            assert _l2rel(x0.cpu(), case["x0"][i]) < st["x"], (prec, key, i)
            assert _l2rel(xp.cpu(), case["x_prev"][i]) < st["x"], (prec, key, i)
            const, _ = closs(x0.clamp(-1, 1))
            scale_y = case["y"].abs().sum()
            assert ((const - case["const"][i]).abs() / scale_y).max() < 5e-2 if prec == "bf16" else 5e-3, (prec, key, i)
            xt = case["x_prev"][i].to(dev)  # teacher forcing: continue from the reference's own x_{t-1}
