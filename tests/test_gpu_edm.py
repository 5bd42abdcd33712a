"""GPU parity of the EDM path (rows N3, G1-EDM, D2, L3): SongUNet / EDM sigma-model executors against the reference's
golden outputs, the fp64 sampler kernels against torch float64, `get_denoise_vector` teacher-forced on the
reference's own x_t of every NFE, and the free-running Heun sampler against the reference's final sample.

Tolerances: networks as in tests/test_gpu_nets.py (max-norm relative 2e-3 tf32 / 2e-2 bf16); sampler kernels
1e-14 relative (float64, same operation order); teacher-forced eps L2-relative 5e-3 (tf32) / 6e-2 (bf16) and
sigma_hat 1e-3 / 1e-2 (operand rounding through the network; there is no discrete time lookup in EDM, so no bucket
flips); free-running final sample (4 Heun steps from sigma 80) PSNR >= 50 dB (tf32) / 35 dB (bf16), peak-to-peak 2."""
import math
import os

import pytest
import torch

from oracle import weights

pytestmark = pytest.mark.gpu
dev = torch.device("cuda:0")
TOL = {"tf32": 2e-3, "fp16": 2e-3, "bf16": 2e-2}


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def _l2rel(a, b):
    return (torch.linalg.vector_norm(a.double() - b.double()) / torch.linalg.vector_norm(b.double()).clamp_min(1e-30)).item()


def _models(prec, name="edm_tiny"):
    from nlc_b200.edm_networks import SigmaModel, SongUNet
    cfg = dict(weights.EDM_CONFIGS[name])
    sg = cfg.pop("sigma")
    m = SongUNet(precision=prec, device=dev, **cfg).load_state_dict(weights.edm_unet_state_dict(**cfg, seed=3))
    s = SigmaModel(dim=sg["dim"], channels=sg["channels"], n_blocks=sg["n_blocks"], precision=prec,
                   device=dev).load_state_dict(weights.edm_sigma_state_dict(**sg, seed=4))
    return cfg, m, s


@pytest.mark.parametrize("prec", ["tf32", "bf16", "fp16"])
def test_networks_golden(golden_dir, prec):
    _, m, s = _models(prec)
    g = torch.load(os.path.join(golden_dir, "nets_edm.pt"), weights_only=True)
    out = m(g["x"].to(dev), g["c_noise"].to(dev))
    feat = m.encode(g["x"].to(dev), g["c_noise"].to(dev))
    assert out.shape == g["out"].shape and feat.shape == g["feat"].shape
    assert _rel(out.cpu(), g["out"]) < TOL[prec]
    assert _rel(feat.cpu(), g["feat"]) < TOL[prec]
    r = s(g["feat"].to(dev))
    assert (r.cpu() - g["r"]).abs().max() < (1e-2 if prec == "bf16" else 1e-3)


def test_fp64_sampler_kernels():
    from nlc_b200 import ops
    from nlc_b200._lib import EDM_PARTS
    g = torch.Generator().manual_seed(2)
    B, shape = 3, (3, 16, 16)
    d = 3 * 16 * 16
    x = (torch.randn(B, *shape, generator=g, dtype=torch.float64) * 7).to(dev)
    F = torch.randn(B, *shape, generator=g).to(dev)
    x32 = torch.empty(B, *shape, device=dev)
    parts = torch.empty(B, EDM_PARTS, device=dev, dtype=torch.float64)
    ops.edm_prepare(x, x32, parts)
    assert torch.equal(x32, x.float())
    assert _rel(parts.sum(1), (x * x).flatten(1).sum(1)) < 1e-14
    cs, co = torch.rand(B, generator=g).to(dev), torch.rand(B, generator=g).to(dev)
    div = (torch.rand(B, generator=g, dtype=torch.float64) + 0.5).to(dev)
    eps, den = torch.empty_like(x), torch.empty_like(x)
    ops.edm_eps(x, x32, F, cs, co, div, eps, den, parts)
    D = (cs.view(B, 1, 1, 1) * x32 + co.view(B, 1, 1, 1) * F).double()
    ref = (x - D) / div.view(B, 1, 1, 1)
    assert torch.equal(den, D) and torch.equal(eps, ref)
    assert _rel(parts.sum(1), (ref * ref).flatten(1).sum(1)) < 1e-14
    e2 = torch.randn(B, *shape, generator=g, dtype=torch.float64).to(dev)
    s1, s2 = div + 0.1, div * 0.7
    nrm = torch.clamp(torch.linalg.vector_norm(eps.flatten(1), dim=1), min=1e-12)
    out = torch.empty_like(x)
    p3 = torch.empty(B, EDM_PARTS, 3, device=dev, dtype=torch.float64)
    ops.edm_mix(eps, nrm, s1, e2, None, s2, 0.3, 0.7, out, p3)
    v1 = (math.sqrt(d) * eps / nrm.view(B, 1, 1, 1)) * s1.view(B, 1, 1, 1)
    refm = 0.3 * v1 + 0.7 * (e2 * s2.view(B, 1, 1, 1))
    assert torch.equal(out, refm)
    sums = p3.sum(1)
    assert _rel(sums[:, 0], (refm * refm).flatten(1).sum(1)) < 1e-13
    assert _rel(sums[:, 2], (refm * v1).flatten(1).sum(1)) < 1e-13
    xn = torch.empty_like(x)
    ops.edm_axpy(x, out, nrm, 1.004, None, s2, xn)
    # (torch's CUDA division by a Python scalar multiplies by the reciprocal; the kernel divides like the CPU reference)
    assert _rel(xn, x + s2.view(B, 1, 1, 1) * ((math.sqrt(d) * out / nrm.view(B, 1, 1, 1)) / 1.004)) < 1e-15


def _experiment(prec):
    from nlc_b200.experiments import EDMImageExperiment
    cfg, m, s = _models(prec)
    Rr = cfg["img_resolution"]
    exp = EDMImageExperiment(m, None, batch_size=2, data_shape=(3, Rr, Rr), seed=1, device=dev, num_timesteps=4,
                             sigma_min=0.002, sigma_max=80)
    exp.set_model(m, s, learn_epsvar=False)
    exp.set_norm_maxmin(0.0, 30.0)
    return exp


@pytest.fixture(scope="module")
def golden_edm(golden_dir):
    return torch.load(os.path.join(golden_dir, "edm_sampler_tiny.pt"), weights_only=False)


CASES = ["pred_partial,pred|00|0|1.0", "base,base|00|0|1.0", "pred,pred_partial|11|1|1.0",
         "pred_sigma,pred_partial3|10|0|None", "pred_partial,pred|01|1|1.004"]


@pytest.mark.parametrize("prec", ["tf32", "bf16", "fp16"])
@pytest.mark.parametrize("key", CASES)
def test_denoise_vector_teacher_forced(golden_edm, prec, key):
    """Every get_denoise_vector call the reference made (7 NFE per case), on the reference's own inputs."""
    style, ne, refine, _ = key.split("|")
    exp = _experiment(prec)
    tol_e, tol_s = (6e-2, 1e-2) if prec == "bf16" else (5e-3, 1e-3)
    for c in golden_edm[key]["calls"]:
        # the dumps hold the noise levels flattened to float64: scalars go back in as 0-d, per-sample ones as [B,1,1,1]
        args = [v.reshape(()) if v.numel() == 1 else v.view(-1, 1, 1, 1).to(dev) for v in (c["sigma_in"], c["sigma_prev_in"])]
        eps, _, s_t, _ = exp.get_denoise_vector(c["xt"].to(dev), args[0], args[1], style=c["style"],
                                                norm_eps=bool(int(ne[0])), refine_prior_sigma=bool(int(refine)))
        assert eps.dtype == torch.float64
        assert _l2rel(eps.cpu(), c["eps"]) < tol_e, (key, c["style"])
        mine, ref = s_t.reshape(-1).cpu().double(), c["sigma_t"]
        assert _l2rel(mine.expand(2), ref.expand(2)) < tol_s, (key, c["style"])


@pytest.mark.parametrize("prec,db", [("tf32", 50.0), ("bf16", 35.0), ("fp16", 45.0)])
@pytest.mark.parametrize("key", ["pred_partial,pred|00|0|1.0", "base,base|00|0|1.0", "pred_partial,pred|01|1|1.004",
                                 "pred_sigma,pred_partial3|10|0|None"])
def test_heun_sampler_free_running(golden_edm, prec, db, key):
    style, ne, refine, es = key.split("|")
    exp = _experiment(prec)
    case = golden_edm[key]
    x = exp.edm_sampler(tuple(case["latents"].shape), latents=case["latents"].to(dev), style=style, norm_eps=ne + "0",
                        refine_prior_sigma=bool(int(refine)), eps_scale=None if es == "None" else float(es))
    assert x.dtype == torch.float64
    mse = ((x.cpu() - case["final"]) ** 2).mean().item()
    psnr = 10 * math.log10(4.0 / max(mse, 1e-30))
    assert psnr >= db, "final-sample PSNR %.1f dB" % psnr


def test_evaluate_edm_is_shardable():
    """Per-sample seeds: rank r of 2 produces exactly the rows the single-process run gives for its batches."""
    exp = _experiment("tf32")
    kw = dict(style="pred_partial,pred", norm_eps="000", microbatch=2)
    full = exp.evaluate_edm(6, **kw)["samples"]
    r0 = exp.evaluate_edm(6, rank=0, world=2, **kw)["samples"]
    r1 = exp.evaluate_edm(6, rank=1, world=2, **kw)["samples"]
    assert torch.equal(torch.cat([r0[:2], r1, r0[2:]]), full)
