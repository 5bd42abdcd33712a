#!/usr/bin/env python
"""Benchmark of the NLC sampling hot path (BASELINE.json metric: DDIM+NLC images/sec).

Workload (c2, BASELINE.json configs[1]): CelebA-64-shaped unet_ddim (ch128, mult 1,2,2,2,4, attn@16) + sigma-model,
ddim_simple_orig eta=0.85, 100 sampling steps from sigma=100, style 'pred' + norm_eps + refine_prior_sigma, clamp
clip, batch 256 per GPU, synthetic seeded weights.  One bench "step" = one full 100-timestep sampling pass of one
batch (= 200 UNet-encoder + 100 decoder + 100 sigma-model evaluations per image).

  value  : images/s with x_T resident in HBM and the result left on the device (kernel-only path)
  e2e    : images/s through the public API (ImageExperiment.denoise_loop): CPU-generator noise, H2D of x_T from
           pinned memory, D2H of the finished images, inside the timed region
  roofline: tcgen05 implicit-GEMM conv kernel, algorithmic FLOPs / CUDA-event time of sampled launches
  cpu_baseline: the oracle port (torch fp32, all host cores) on a bounded sample of the same workload

The default run's ONE JSON line also carries (sub-records, each measured by the same timing code):
  adm256 : the north-star target config c5 (ADM-256 + colourisation, batch 64 per GPU, 10 timesteps per pass): value,
           ms_per_timestep, roofline, e2e - so the driver's 1->8 scaling run measures the target config too
  tf32   : c2 in the tf32 operand mode (what cuDNN does for the reference's default GPU run)
  bf16   : c2 in the other 16-bit operand mode (the headline mode is fp16: HEADLINE_PRECISION)
  strong : (N > 1) c2 with the GLOBAL batch 256 sharded over the N GPUs (configs[1]'s wording), exact batch-global
           decisions via the per-step all-reduce
  kernel_to_beat : the reference's networks through PyTorch eager / cuDNN (TF32 convolutions, its default GPU path) on
           the same GPU, one NLC timestep at the workload's batch (the oracle's functional networks moved to the GPU;
           baseline leg only, rank 0)
`--no-extras` keeps only the headline record.

`--impl reference` times the oracle port (the reference is pure Python and cannot travel to the GPU box) on the
host cores for the same config.  N>1: one process per GPU under torchrun, batch sharded, NCCL all-gather of the
finished images each pass; value = all ranks' images / max-over-ranks device time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# Workloads.  c2 (BASELINE.json configs[1]) is the default and the one the driver's bench line is quoted on; c4 / c5 are
# the ADM-256 restoration configs (configs[3], configs[4]), run with `--workload c4|c5|c5cs` for the ADM-256 numbers of
# the north-star (tensor-pipe utilisation of the conv/attention GEMMs at batch 64 per GPU).  GFLOP per NFE: SURVEY §8(d).
_NLC = dict(sampler="ddim_simple_orig", eta=0.85, start_sigma=100.0, style="pred", norm_eps=True, refine=True,
            norm_min=-2.0, norm_max=110.0, sigma_pred_threshold=960)
WORKLOADS = {
    "c2": dict(_NLC, name="c2", arch="ddim", R=64, steps=100, clip="clamp", batch=256, gflop_per_nfe=60.57,
               constraint=None, learn_epsvar=False, sampler_var="none",
               label="c2: CelebA-64 unet_ddim + sigma-model, ddim_simple_orig eta 0.85, %d steps, NLC pred",
               metric="DDIM+NLC images/sec (CelebA-64 unet_ddim, 100 steps)"),
    "c3": dict(name="edm64", arch="edm", R=64, steps=18, batch=512, gflop_per_nfe=108.07, nfe_per_pass=35,
               style="pred_partial,pred", norm_eps="000", refine=False, constraint=None, norm_min=-2.0, norm_max=110.0,
               label="c3: EDM SongUNet-64 (DDPM++) + sigma-model, Heun sampler, %d steps = 35 NFE, NLC pred_partial,pred",
               metric="EDM Heun+NLC images/sec (SongUNet-64, 35 NFE)"),
    "c4": dict(_NLC, name="adm256", arch="adm", R=256, steps=10, clip="dynamic", batch=32, gflop_per_nfe=2823.91,
               constraint=("sr_averagepooling", 4.0), learn_epsvar=True, sampler_var="learned",
               label="c4: ADM-256 UNet + sigma-model, DDNM SR x4 (svd projection), ddim_simple_orig eta 0.85, dynamic "
                     "clip, %d steps, NLC pred",
               metric="DDIM+NLC images/sec (ADM-256, DDNM SR x4)"),
    "c5": dict(_NLC, name="adm256", arch="adm", R=256, steps=10, clip="dynamic", batch=64, gflop_per_nfe=2823.91,
               constraint=("colorization", 1.0), learn_epsvar=True, sampler_var="learned",
               label="c5: ADM-256 UNet + sigma-model, DDNM colourisation (svd projection), ddim_simple_orig eta 0.85, "
                     "dynamic clip, %d steps, NLC pred",
               metric="DDIM+NLC images/sec (ADM-256, DDNM colourisation)"),
    "c5cs": dict(_NLC, name="adm256", arch="adm", R=256, steps=10, clip="dynamic", batch=64, gflop_per_nfe=2823.91,
                 constraint=("cs_walshhadamard", 4.0), learn_epsvar=True, sampler_var="learned",
                 label="c5: ADM-256 UNet + sigma-model, DDNM Walsh-Hadamard CS x4 (svd projection), ddim_simple_orig eta "
                       "0.85, dynamic clip, %d steps, NLC pred",
                 metric="DDIM+NLC images/sec (ADM-256, DDNM WH-CS x4)"),
    # the DDNM+ sampler of functions/svd_ddnm.py (noisy measurements, one UNet forward per
    # step, no sigma-model) on the c4 network and operator
    "c4p": dict(name="adm256", arch="adm", loop="ddnm_plus", R=256, steps=10, batch=32, gflop_per_nfe=2239.67, eta=0.85,
                sigma_y=0.1, constraint=("sr_averagepooling", 4.0),
                label="c4+: ADM-256 UNet, DDNM+ SR x4 with measurement noise 0.1 (functions/svd_ddnm.py "
                      "ddnm_plus_diffusion, eta 0.85), %d steps",
                metric="DDNM+ images/sec (ADM-256, SR x4, noisy measurements)"),
}
CFG = dict(WORKLOADS["c2"])
ADM_KEYS = ("image_size", "model_channels", "out_channels", "num_res_blocks", "attention_resolutions", "channel_mult",
            "num_heads", "num_head_channels", "use_scale_shift_norm", "resblock_updown", "use_new_attention_order")
# The operand mode of the driver's line: fp16 - the 16-bit mode (same tcgen05 kind::f16 rate as bf16) that meets the
# north-star's 45 dB gate with margin (63 dB with the reference's time buckets against 46 dB for bf16),
# and the reference's own reduced-precision mode (src/fp16_util.py).  bf16 and tf32 are reported as sub-records.
HEADLINE_PRECISION = "fp16"
CPU_SAMPLE = {"ddim": (6, 32), "adm": (1, 2), "edm": (3, 16)}  # (timesteps, batch) of the bounded CPU sample: ~10-20 s of host work


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d.get("bf16_tflops_sustained", 1399.0), d.get("hbm_gbs", 6527.8), "measured (MEASURED_PEAKS.json, sustained)"
    return 1400.0, 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (profiling recipe's clocks line)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.rows.append(parts)
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        mhz = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": mhz[len(mhz) // 2] if mhz else None,
                "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None, "reasons": reasons}


def cpu_port_rate(n_steps, batch, threads):
    """images/s of the oracle port on the host: `n_steps` NLC sampling steps at `batch`, extrapolated to the
    workload's step count.  The constraint projection is left out of the CPU sample (it is <1 % of a step)."""
    from oracle import sampler as S, weights
    torch.set_num_threads(threads)
    if CFG["arch"] == "edm":
        # `n_steps` Heun steps (2 NFE each, the last one Euler only) of the reference's sigma ladder at `batch`
        from oracle import edm_net, sampler_edm
        cfg = dict(weights.EDM_CONFIGS[CFG["name"]])
        sg = cfg.pop("sigma")
        sd = weights.edm_unet_state_dict(**cfg, seed=3)
        ssd = weights.edm_sigma_state_dict(**sg, seed=4)
        d = 3 * CFG["R"] ** 2
        o = sampler_edm.EDM(lambda x, c: edm_net.unet_forward(sd, x, c, cfg), lambda x, c: edm_net.unet_encode(sd, x, c, cfg),
                            lambda f: edm_net.sigma_forward(ssd, f), d, norm_min=CFG["norm_min"] / d ** 0.5,
                            norm_max=CFG["norm_max"] / d ** 0.5)
        lat = torch.randn(batch, 3, CFG["R"], CFG["R"], generator=torch.Generator().manual_seed(0))
        with torch.no_grad():
            t0 = time.perf_counter()
            o.sample(lat, n_steps, style=CFG["style"], norm_eps=CFG["norm_eps"], refine=CFG["refine"])
            dt = time.perf_counter() - t0
        nfe = 2 * n_steps - 1
        return batch / (dt / nfe * CFG["nfe_per_pass"]), dt
    if CFG.get("loop") == "ddnm_plus":
        # `n_steps` reverse steps (one network call each) of the oracle's DDNM+ loop at `batch`
        from oracle import adm_net, ddnm, operators as O
        cfg = dict(weights.ADM_CONFIGS[CFG["name"]])
        cfg.pop("sigma")
        sd = weights.adm_unet_state_dict(**cfg, seed=3)
        R = CFG["R"]
        op = O.SuperResolution(3, R, int(CFG["constraint"][1]))
        g = torch.Generator().manual_seed(0)
        y = op.A(torch.rand(batch, 3 * R * R, generator=g) * 2 - 1)
        x = torch.randn(batch, 3, R, R, generator=g)
        with torch.no_grad():
            t0 = time.perf_counter()
            ddnm.run(x, lambda z, t: adm_net.unet_forward(sd, z, t, cfg), torch.linspace(1e-4, 2e-2, 1000), CFG["eta"], op,
                     y, CFG["sigma_y"], T_sampling=n_steps, noise_fn=lambda like: torch.randn(like.shape, generator=g))
            dt = time.perf_counter() - t0
        return batch / (dt / n_steps * CFG["steps"]), dt
    if CFG["arch"] == "adm":
        from oracle import adm_net
        cfg = dict(weights.ADM_CONFIGS[CFG["name"]])
        sg = cfg.pop("sigma")
        sd = weights.adm_unet_state_dict(**cfg, seed=3)
        ssd = weights.adm_sigma_state_dict(**sg, seed=4)
        fwd = lambda z, t: adm_net.unet_forward(sd, z, t, cfg)
        enc = lambda z, t: adm_net.unet_encode(sd, z, t, cfg)
        sg_fn = lambda f: adm_net.sigma_forward(ssd, f, cfg)
    else:
        from oracle import ddim_net
        cfg = weights.CONFIGS[CFG["name"]]
        sd = weights.ddim_unet_state_dict(**cfg["unet"], seed=3)
        ssd = weights.ddim_sigma_state_dict(**cfg["sigma"], seed=4)
        fwd = lambda z, t: ddim_net.unet_forward(sd, z, t)
        enc = lambda z, t: ddim_net.unet_encode(sd, z, t)
        sg_fn = lambda f: ddim_net.sigma_forward(ssd, f)
    tab = S.Tables()
    ts, sig, mvc = tab.ddim_schedule(CFG["start_sigma"], None, CFG["steps"])
    d = 3 * CFG["R"] ** 2
    g = torch.Generator().manual_seed(0)
    shape = (batch, 3, CFG["R"], CFG["R"])
    xT = torch.randn(shape, generator=g) / (1 / (sig[0] ** 2 + 1)).sqrt()
    noises = [torch.randn(shape, generator=g) for _ in range(n_steps + 1)]
    first = next(i for i, t in enumerate(ts.tolist()) if t <= CFG["sigma_pred_threshold"])  # NLC-active steps
    first = min(first, len(ts) - 2 - n_steps)
    args = dict(kind=CFG["sampler"], eta=CFG["eta"], style=CFG["style"], norm_eps=CFG["norm_eps"], refine=CFG["refine"],
                norm_min=CFG["norm_min"] / d ** 0.5, norm_max=CFG["norm_max"] / d ** 0.5, clip=CFG["clip"],
                sampler_var=CFG["sampler_var"], learn_epsvar=CFG["learn_epsvar"])
    with torch.no_grad():
        if CFG["arch"] != "adm":  # warm-up (the ADM sample is a single 10+ s step: no separate warm-up)
            S.denoise_loop(tab, ts[first:first + 2].tolist(), sig[first:first + 2], mvc, fwd, enc, sg_fn, xT,
                           noises=noises, **args)
        t0 = time.perf_counter()
        S.denoise_loop(tab, ts[first:first + n_steps + 1].tolist(), sig[first:first + n_steps + 1], mvc, fwd, enc, sg_fn,
                       xT, noises=noises, **args)
        dt = time.perf_counter() - t0
    return batch / (dt / n_steps * CFG["steps"]), dt


def _unit():
    """What one sampled CPU step is, for the `sample` text."""
    if CFG.get("loop") == "ddnm_plus":
        return "DDNM+ reverse steps (one network call each)"
    return "Heun steps" if CFG["arch"] == "edm" else "NLC timesteps"


def _select(args):
    CFG.clear()
    CFG.update(WORKLOADS[args.workload])
    if args.timesteps:
        CFG["steps"] = args.timesteps
    if args.batch:
        CFG["batch"] = args.batch
    CFG["label"] = CFG["label"] % CFG["steps"]


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    n_steps, batch = CPU_SAMPLE[CFG["arch"]]
    per_step = []
    for _ in range(args.warmup if CFG["arch"] != "adm" else 0):
        cpu_port_rate(1, 4, threads)
    for _ in range(args.steps):
        per_step.append(cpu_port_rate(n_steps, batch, threads))
    rate = sum(r for r, _ in per_step) / len(per_step)
    line = {
        "impl": "reference", "metric": CFG["metric"], "value": rate,
        "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1000.0 * CFG["batch"] / rate, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32", "data": "synthetic",
        "config": {"workload": CFG["label"], "per_gpu_batch": CFG["batch"]},
        "cpu_baseline": {"value": rate, "unit": "images/s", "cores": threads, "kind": "port",
                         "sample": ("each step = %d " + _unit() + " at batch %d of the workload on the oracle port "
                                    "(torch fp32, %d threads, %.1f s per step), extrapolated linearly to %d timesteps")
                                   % (n_steps, batch, threads, sum(d for _, d in per_step) / len(per_step),
                                      CFG["steps"])},
        "e2e": {"value": rate, "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def build_models(precision, dev):
    """The workload's UNet + sigma-model executors with seeded synthetic weights (nlc_b200.synthetic_weights: random-init
    state_dicts of the named architectures; nothing under oracle/ is imported on this arm)."""
    from nlc_b200 import synthetic_weights as weights
    if CFG["arch"] == "edm":
        from nlc_b200.edm_networks import SigmaModel, SongUNet
        cfg = dict(weights.EDM_CONFIGS[CFG["name"]])
        sg = cfg.pop("sigma")
        model = SongUNet(precision=precision, device=dev, **cfg).load_state_dict(
            weights.edm_unet_state_dict(**cfg, seed=3))
        sigma_model = SigmaModel(dim=sg["dim"], channels=sg["channels"], n_blocks=sg["n_blocks"], precision=precision,
                                 device=dev).load_state_dict(weights.edm_sigma_state_dict(**sg, seed=4))
        return model, sigma_model
    if CFG["arch"] == "adm":
        from nlc_b200.unet_adm import SigmaModel, UNetModel
        cfg = dict(weights.ADM_CONFIGS[CFG["name"]])
        sg = cfg.pop("sigma")
        model = UNetModel(in_channels=3, precision=precision, device=dev, **{k: cfg[k] for k in ADM_KEYS}).load_state_dict(
            weights.adm_unet_state_dict(**cfg, seed=3))
        sigma_model = SigmaModel(dim=sg["dim"], channels=sg["channels"], n_blocks=sg["n_blocks"],
                                 num_heads=cfg["num_heads"], num_head_channels=cfg["num_head_channels"],
                                 precision=precision, device=dev).load_state_dict(
            weights.adm_sigma_state_dict(**sg, seed=4))
        return model, sigma_model
    from nlc_b200.unet_ddim import SigmaModel, UNetModel
    cfg = weights.CONFIGS[CFG["name"]]
    model = UNetModel(**cfg["unet"], precision=precision, device=dev).load_state_dict(
        weights.ddim_unet_state_dict(**cfg["unet"], seed=3))
    sigma_model = SigmaModel(**cfg["sigma"], precision=precision, device=dev).load_state_dict(
        weights.ddim_sigma_state_dict(**cfg["sigma"], seed=4))
    return model, sigma_model


PROFILE_TRAFFIC = os.path.join(ROOT, "profiles", "ncu_traffic.json")


def measured_traffic(workload, batch, precision):
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the workload's dominant conv_tc shape, from the
    committed ncu --set full summary (profiles/ncu_traffic.json, written by scripts/ncu_summary.py from the .ncu-rep of the
    final binary); None when no capture exists for this (workload, batch, precision)."""
    try:
        with open(PROFILE_TRAFFIC) as f:
            t = json.load(f)
        e = t.get("%s|%d|%s" % (workload, batch, precision))
        return (e["dram_bytes"], e.get("source")) if e else (None, None)
    except Exception:
        return None, None


class DistCtx:
    def __init__(self):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            # NCCL prints its version banner on STDOUT at NCCL_DEBUG=VERSION and =WARN; the contract is ONE JSON line
            # there.  Quieten those two levels and, whatever the level, send everything NCCL prints while the
            # communicator comes up (eager init with device_id + one barrier) to stderr.
            if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
                os.environ["NCCL_DEBUG"] = "NONE"
            sys.stdout.flush()
            saved_stdout = os.dup(1)
            os.dup2(2, 1)
            try:
                dist.init_process_group("nccl", device_id=self.dev)
                dist.barrier()
                torch.cuda.synchronize()
            finally:
                sys.stdout.flush()
                os.dup2(saved_stdout, 1)
                os.close(saved_stdout)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(self, v):
        if self.world == 1:
            return v
        t = torch.tensor([v], device=self.dev, dtype=torch.float64)
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())


def measure(dc, precision, steps, warmup, use_graph=True, do_e2e=True, exact_global=False, scaling="weak"):
    """Time the workload selected in CFG on this rank's GPU: `warmup` untimed passes, then exactly `steps` passes between
    barrier + synchronize, CUDA events on the launching stream, max over ranks.  Returns the record's fields."""
    from nlc_b200 import constraint_functions as CF, ops
    from nlc_b200.experiments import ImageExperiment
    from nlc_b200.schedulers import get_sampler

    import gc
    gc.collect()  # (the previous measurement's networks and plans: tens of GB of device buffers)
    torch.cuda.empty_cache()
    dist, world, rank, dev = dc.dist, dc.world, dc.rank, dc.dev
    B, R = CFG["batch"], CFG["R"]
    model, sigma_model = build_models(precision, dev)
    shape = (B, 3, R, R)
    g_dev = torch.Generator(device=dev).manual_seed(99 + rank)
    y_bytes = 0
    gathered = None
    conv_samples = []
    # conv kernels are timed on every 10th (5th) timestep of the timed passes: those timesteps are launched eagerly with one
    # CUDA-event pair per conv launch, the others replay the captured graph
    every = 10 if CFG["steps"] >= 20 else 5
    state = dict(active=False)

    def sampled(ind):
        return ind % every == every // 2

    def hook(ind, _):  # EDM / DDNM+ loops: arm the per-launch timer for the next step
        ops.STATS.conv_timer = conv_samples if sampled(ind + 1) and state["active"] else None

    def graph_skip(ind):  # denoise_loop: the sampled timesteps run eagerly with the per-launch timer, the rest replay
        on = sampled(ind) and state["active"]
        ops.STATS.conv_timer = conv_samples if on else None
        return on

    def step_done(ind, _):
        ops.STATS.conv_timer = None

    graphed = False
    if CFG.get("loop") == "ddnm_plus":
        import types
        from nlc_b200 import svd_ddnm
        task, scale = CFG["constraint"]
        A_funcs = CF.svd_constraint(task, fn_scale=scale, device=dev, image_size=R, channels=3)
        x_true = torch.rand(shape, generator=g_dev, device=dev) * 2 - 1
        y = A_funcs.A(x_true)
        y = y + CFG["sigma_y"] * torch.randn(y.shape, generator=g_dev, device=dev)
        y_bytes = y.numel() * 4
        betas = torch.linspace(1e-4, 2e-2, 1000, device=dev)
        ddnm_cfg = types.SimpleNamespace(
            diffusion=types.SimpleNamespace(num_diffusion_timesteps=1000),
            time_travel=types.SimpleNamespace(T_sampling=CFG["steps"], travel_length=1, travel_repeat=1))
        lat_dev = torch.randn(shape, generator=g_dev, device=dev)
        out_dtype, x_scale = torch.float32, 1.0
        calls = [0]

        def net(x, t):
            hook(calls[0] - 1, None)
            calls[0] += 1
            return model(x, t)

        if world > 1:
            gathered = [torch.empty(shape, device=dev) for _ in range(world)]

        def sample(x_in, to_cpu, with_hook):
            calls[0] = 0
            xs, x0s = svd_ddnm.ddnm_plus_diffusion(x_in, net if with_hook else model, betas, CFG["eta"], A_funcs, y,
                                                   CFG["sigma_y"], config=ddnm_cfg, to_cpu=to_cpu)
            return x0s[0]
    elif CFG["arch"] == "edm":
        from nlc_b200.experiments import EDMImageExperiment
        exp = EDMImageExperiment(model, None, batch_size=B, data_shape=(3, R, R), seed=1234 + rank, device=dev,
                                 num_timesteps=CFG["steps"])
        exp.set_model(model, sigma_model, learn_epsvar=False)
        exp.set_norm_maxmin(CFG["norm_min"], CFG["norm_max"])
        edm_kw = dict(style=CFG["style"], norm_eps=CFG["norm_eps"], refine_prior_sigma=CFG["refine"])
        lat_dev = torch.randn(shape, generator=g_dev, device=dev)
        out_dtype, x_scale = torch.float64, 1.0
        if world > 1:
            gathered = [torch.empty(shape, device=dev, dtype=torch.float64) for _ in range(world)]

        def sample(x_in, to_cpu, with_hook):
            out = exp.edm_sampler(shape, latents=x_in, step_hook=(lambda i, d: hook(i, d)) if with_hook else None,
                                  **edm_kw)
            return out.cpu() if to_cpu else out
    else:
        graphed = use_graph
        sch = get_sampler(CFG["sampler"], 1000, CFG["steps"], start_sigma=CFG["start_sigma"], eta=CFG["eta"],
                          sampler_var=CFG["sampler_var"]).to(dev)
        exp = ImageExperiment(model, sch, batch_size=B, data_shape=(3, R, R), seed=1234 + rank, device=dev)
        exp.set_model(model, sigma_model, learn_epsvar=CFG["learn_epsvar"])
        exp.set_norm_maxmin(CFG["norm_min"], CFG["norm_max"])
        exp.set_clip_fn(CFG["clip"])
        loop_kw = dict(style=CFG["style"], norm_eps=CFG["norm_eps"], refine_prior_sigma=CFG["refine"], return_log=False,
                       chunk_size=1, sigma_pred_threshold=CFG["sigma_pred_threshold"], graph=use_graph,
                       exact_global=exact_global and world > 1)
        if CFG["constraint"] is not None:
            # DDNM restoration: synthetic ground truth x ~ U(-1,1), measurement y = A x, projection + loss every step
            # (image_sample.py:636-665); the CS permutation is drawn once on the CPU with a fixed seed
            from functools import partial
            task, scale = CFG["constraint"]
            con = CF.get_constraint_function(task, constraint_scale=scale, device=dev, image_size=R, channels=3,
                                             perm=torch.randperm(R * R, generator=torch.Generator().manual_seed(7)))
            x_true = torch.rand(shape, generator=g_dev, device=dev) * 2 - 1
            y = con.transform(x_true)
            y_bytes = y.numel() * 4
            loop_kw.update(constrain_fn=partial(con.constraint_fn, y=y, lambda_t=con.lr),
                           constrain_loss=partial(con.loss, y=y))
        x_scale = (float(sch.sampling_sigmas[0]) ** 2 + 1) ** 0.5
        lat_dev = torch.randn(shape, generator=g_dev, device=dev) * x_scale
        out_dtype = torch.float32
        if world > 1:
            gathered = [torch.empty(shape, device=dev) for _ in range(world)]

        def sample(x_in, to_cpu, with_hook):
            out, _ = exp.denoise_loop(shape=shape, xT=x_in, to_cpu=to_cpu,
                                      graph_skip=graph_skip if with_hook else None,
                                      step_hook=step_done if with_hook else None, **loop_kw)
            return out

    def one_pass_device():
        out = sample(lat_dev, False, True)
        if world > 1:
            dist.all_gather(gathered, out.contiguous())
        return out

    # ---------------------------------------------------------------- device-resident timing
    for _ in range(max(warmup, 1)):
        one_pass_device()
    dc.barrier()
    clocks = ClockSampler(dc.local)
    clocks.start()
    ops.STATS.launches = 0
    ops.STATS.graph_replays = 0
    state["active"] = True
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        one_pass_device()
    e1.record()
    dc.barrier()
    state["active"] = False
    ops.STATS.conv_timer = None
    launches, replays = ops.STATS.launches, ops.STATS.graph_replays
    clocks.stop_flag = True
    ms = dc.max_over_ranks(e0.elapsed_time(e1))
    ms_per_step = ms / steps
    value = B * world / (ms_per_step / 1000.0)

    # ---------------------------------------------------------------- end-to-end through the public API
    e2e = None
    if do_e2e:
        pinned = torch.empty(shape, dtype=torch.float32).pin_memory()
        gen = torch.Generator().manual_seed(4321 + rank)

        def one_pass_e2e():
            # the reference draws the initial noise on the host side (src/experiments.py:268; per-sample generators for
            # EDM): pinned staging, H2D, the whole sampling loop through the public API, D2H of the finished images
            torch.randn(shape, generator=gen, out=pinned)
            x = pinned.to(dev, non_blocking=True)
            if x_scale != 1.0:
                x = x * x_scale
            return sample(x, True, False)

        one_pass_e2e()
        dc.barrier()
        t0 = time.perf_counter()
        n_e2e = max(1, min(steps, 3))
        for _ in range(n_e2e):
            one_pass_e2e()
        dc.barrier()
        e2e_s = dc.max_over_ranks((time.perf_counter() - t0) / n_e2e)
        nbytes = B * 3 * R * R * 4
        out_bytes = B * 3 * R * R * (8 if out_dtype == torch.float64 else 4)
        e2e = {"value": B * world / e2e_s, "unit": "images/s", "h2d_bytes_per_step": nbytes + y_bytes,
               "d2h_bytes_per_step": out_bytes}

    # ---------------------------------------------------------------- roofline of the dominant kernel
    tf_peak, hbm_peak, peak_src = peaks()
    conv_ms = sum(a.elapsed_time(b) for _, a, b in conv_samples)
    conv_fl = sum(f for f, _, _ in conv_samples)
    achieved = conv_fl / (conv_ms / 1000.0) / 1e12 if conv_ms > 0 else 0.0
    nfe_per_pass = CFG.get("nfe_per_pass", CFG["steps"])
    nfe_flops = CFG["gflop_per_nfe"] * 1e9 * B * nfe_per_pass
    sampled_timesteps = max(1, steps * sum(1 for i in range(CFG["steps"]) if sampled(i)))
    traffic, traffic_src = measured_traffic(CFG["name"], B, precision)
    roofline = {"bound": "tensor", "kernel": "nlc::conv_tc_kernel (tcgen05 implicit GEMM, %s)" % precision,
                "achieved": achieved, "peak": tf_peak, "unit": "TFLOP/s", "frac": achieved / tf_peak,
                "peak_source": peak_src,
                # dram__bytes_read + dram__bytes_write of one launch of the step's dominant conv shape, read from the
                # committed ncu --set full summary of the final binary (None: no capture for this batch / mode)
                "traffic": traffic, "traffic_source": traffic_src,
                "launches_sampled": len(conv_samples),
                # share of the step spent in the conv kernel, from the sampled (eagerly launched) timesteps
                "conv_share_of_step": (conv_ms / sampled_timesteps) / (ms_per_step / CFG["steps"]) if conv_ms > 0 else None,
                "whole_step_tflops": nfe_flops / (ms_per_step / 1000.0) / 1e12,
                "whole_step_frac": nfe_flops / (ms_per_step / 1000.0) / 1e12 / tf_peak}
    rec = {
        "metric": CFG["metric"], "value": value, "unit": "images/s", "n_gpus": world, "steps": steps, "warmup": warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": precision, "data": "synthetic",
        "config": {"workload": CFG["label"], "per_gpu_batch": B, "global_batch": B * world,
                   "step": "one full %d-timestep sampling pass of one batch" % CFG["steps"],
                   "l2": "activations per pass (GBs) exceed the 126 MB L2; no explicit flush",
                   "parallelism": "dp%d (batch sharded, NCCL all-gather of finished images%s)" % (
                       world, "; per-step all-reduce of the batch-global loop decisions" if exact_global and world > 1 else ""),
                   "cuda_graph": bool(graphed)},
        "nfe_per_s": value * nfe_per_pass,
        "ms_per_timestep": ms_per_step / CFG["steps"],
        "clocks": clocks.summary(),
        "e2e": e2e,
        "gpu_launches": launches,
        "graph_replays": replays,
        "roofline": roofline,
    }
    del model, sigma_model
    torch.cuda.empty_cache()
    return rec


def kernel_to_beat(dev, reps=3):
    """The reference's networks through PyTorch eager / cuDNN on this GPU (its own GPU path: fp32 tensors, TF32 allowed for
    convolutions): one NLC timestep = UNet encode + sigma-model + UNet forward at the workload's batch, via the oracle's
    functional networks (pinned torch.equal to the reference on the CPU) moved to the device.  Baseline leg only."""
    from oracle import weights
    torch.backends.cudnn.benchmark = True
    torch.backends.cudnn.allow_tf32 = True
    B, R = CFG["batch"], CFG["R"]
    if CFG["arch"] == "adm":
        from oracle import adm_net
        B = min(B, 16)  # torch eager needs 161 GB at batch 32 (profiles/r01g_torch_eager_baseline.log)
        cfg = dict(weights.ADM_CONFIGS[CFG["name"]])
        sg = cfg.pop("sigma")
        sd = {k: v.to(dev) for k, v in weights.adm_unet_state_dict(**cfg, seed=3).items()}
        ssd = {k: v.to(dev) for k, v in weights.adm_sigma_state_dict(**sg, seed=4).items()}
        enc, fwd = (lambda x, t: adm_net.unet_encode(sd, x, t, cfg)), (lambda x, t: adm_net.unet_forward(sd, x, t, cfg))
        sig = lambda f: adm_net.sigma_forward(ssd, f, cfg)
    elif CFG["arch"] == "ddim":
        from oracle import ddim_net
        cfg = weights.CONFIGS[CFG["name"]]
        sd = {k: v.to(dev) for k, v in weights.ddim_unet_state_dict(**cfg["unet"], seed=3).items()}
        ssd = {k: v.to(dev) for k, v in weights.ddim_sigma_state_dict(**cfg["sigma"], seed=4).items()}
        enc, fwd = (lambda x, t: ddim_net.unet_encode(sd, x, t)), (lambda x, t: ddim_net.unet_forward(sd, x, t))
        sig = lambda f: ddim_net.sigma_forward(ssd, f)
    else:
        return None
    x = torch.randn(B, 3, R, R, device=dev)
    t = torch.full((B,), 500.0, device=dev)

    def step():
        r = sig(enc(x, t))
        return fwd(x * (1 + r).reshape(-1, 1, 1, 1), t)

    out = {}
    for mode, ctx in (("tf32", None), ("fp16_autocast", torch.autocast("cuda", dtype=torch.float16))):
        with torch.no_grad():
            if ctx is not None:
                ctx.__enter__()
            try:
                for _ in range(2):
                    step()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    step()
                e1.record()
                torch.cuda.synchronize()
            finally:
                if ctx is not None:
                    ctx.__exit__(None, None, None)
        ms = e0.elapsed_time(e1) / reps
        out[mode] = {"ms_per_timestep": ms, "value": B / (ms / 1000.0 * CFG["steps"]), "unit": "images/s"}
    out["batch"] = B
    out["what"] = ("torch %s eager + cuDNN on this GPU, the reference's networks (oracle functional form): encode + "
                   "sigma-model + forward per NLC timestep, x %d timesteps; tf32 = the reference's default GPU path, "
                   "fp16_autocast = torch.autocast(float16), not a reference mode" % (torch.__version__, CFG["steps"]))
    del sd, ssd
    torch.cuda.empty_cache()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="nlc", choices=["nlc", "reference"])
    ap.add_argument("--workload", default="c2", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="images per GPU (default: the workload's)")
    ap.add_argument("--precision", default=HEADLINE_PRECISION, choices=["bf16", "fp16", "tf32", "fp32"])
    ap.add_argument("--timesteps", type=int, default=0, help="sampling steps per pass (default: the workload's)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="strong: the workload's batch is the GLOBAL batch, sharded over the GPUs")
    ap.add_argument("--no-graph", action="store_true", help="launch every timestep eagerly (no CUDA-graph replay)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="headline record only (no adm256 / tf32 / strong / "
                                                             "kernel_to_beat sub-records)")
    args = ap.parse_args()
    _select(args)
    if args.impl == "reference":
        return run_reference(args)

    dc = DistCtx()
    if args.scaling == "strong":
        assert CFG["batch"] % dc.world == 0, "the global batch must divide over the GPUs"
        CFG["batch"] //= dc.world
    line = measure(dc, args.precision, args.steps, args.warmup, use_graph=not args.no_graph,
                   exact_global=args.scaling == "strong", scaling=args.scaling)
    extras = args.workload == "c2" and not args.no_extras and not args.batch and not args.timesteps \
        and args.scaling == "weak"
    if extras:
        # tf32 operand mode of the same workload (one timed pass)
        sub = measure(dc, "tf32", 1, 1, use_graph=not args.no_graph, do_e2e=False)
        line["tf32"] = {k: sub[k] for k in ("value", "unit", "ms_per_timestep", "dtype")}
        line["tf32"]["roofline_frac_of_bf16_peak"] = sub["roofline"]["frac"]
        # ... and the other 16-bit operand mode (bf16 when the headline is fp16, and the other way round)
        other = "bf16" if args.precision != "bf16" else "fp16"
        sub = measure(dc, other, 1, 1, use_graph=not args.no_graph, do_e2e=False)
        line[other] = {k: sub[k] for k in ("value", "unit", "ms_per_timestep", "dtype")}
        line[other]["roofline_frac"] = sub["roofline"]["frac"]
        if dc.world > 1:
            # configs[1] as worded: global batch 256 sharded over the GPUs (strong scaling), exact batch-global decisions
            CFG["batch"] = WORKLOADS["c2"]["batch"] // dc.world
            sub = measure(dc, args.precision, 2, 2, use_graph=not args.no_graph, do_e2e=False, exact_global=True,
                          scaling="strong")
            line["strong"] = {k: sub[k] for k in ("value", "unit", "ms_per_timestep", "dtype", "scaling", "gpu_launches",
                                                  "graph_replays")}
            line["strong"]["per_gpu_batch"], line["strong"]["global_batch"] = CFG["batch"], WORKLOADS["c2"]["batch"]
            # ... and the same sharded run launched eagerly: what the CUDA graph buys at 32 images per GPU
            sub = measure(dc, args.precision, 1, 1, use_graph=False, do_e2e=False, exact_global=True, scaling="strong")
            line["strong"]["eager_value"] = sub["value"]
        # the north-star target config: ADM-256, batch 64 per GPU (c5), weak scaling
        CFG.clear()
        CFG.update(WORKLOADS["c5"])
        CFG["label"] = CFG["label"] % CFG["steps"]
        sub = measure(dc, args.precision, 2, 2, use_graph=not args.no_graph)
        line["adm256"] = {k: sub[k] for k in ("metric", "value", "unit", "n_gpus", "ms_per_step", "ms_per_timestep", "dtype",
                                              "scaling", "config", "e2e", "roofline", "gpu_launches", "graph_replays",
                                              "nfe_per_s", "clocks")}
        CFG.clear()
        CFG.update(WORKLOADS["c2"])
        CFG["label"] = CFG["label"] % CFG["steps"]

    if dc.rank != 0:
        if dc.world > 1:
            dc.dist.destroy_process_group()
        return
    if dc.world > 1:
        dc.dist.destroy_process_group()  # the baseline legs below are rank 0's alone

    if not args.no_extras:
        try:
            with torch.no_grad():
                line["kernel_to_beat"] = kernel_to_beat(dc.dev)
        except Exception as e:  # a baseline leg must never cost the headline line
            line["kernel_to_beat"] = {"unavailable": "%s: %s" % (type(e).__name__, e)}
    cpu = None
    if not args.no_cpu_baseline:
        threads = os.cpu_count() or 1
        n_steps, batch = CPU_SAMPLE[CFG["arch"]]
        rate, dt = cpu_port_rate(n_steps, batch, threads)
        cpu = {"value": rate, "unit": "images/s", "cores": threads, "kind": "port",
               "sample": ("%d " + _unit() + " at batch %d of the workload on the oracle port (torch fp32, %d threads, "
                          "%.1f s), extrapolated linearly to %d timesteps") % (n_steps, batch, threads, dt, CFG["steps"])}
    line["cpu_baseline"] = cpu
    print(json.dumps(line), flush=True)


if __name__ == "__main__":
    main()
