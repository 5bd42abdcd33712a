"""The oracle (CPU restatement under oracle/) against golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  Bit-exact where the arithmetic is the same sequence of torch ops."""
import os

import pytest
import torch

from oracle import ddim_net, operators as O, sampler as S, weights

torch.set_num_threads(4)


def load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=True)


@pytest.fixture(scope="module")
def tiny():
    cfg = weights.CONFIGS["tiny"]
    return (weights.ddim_unet_state_dict(**cfg["unet"], seed=3), weights.ddim_sigma_state_dict(**cfg["sigma"], seed=4))


def test_networks(golden_dir, tiny):
    sd, ssd = tiny
    g = load(golden_dir, "nets_tiny.pt")
    with torch.no_grad():
        out, feat = ddim_net.unet_forward(sd, g["x"], g["t"], return_feat=True)
        enc = ddim_net.unet_encode(sd, g["x"], g["t"])
        r = ddim_net.sigma_forward(ssd, feat)
    assert torch.equal(out, g["out"]) and torch.equal(feat, g["feat"]) and torch.equal(enc, g["feat"])
    assert torch.equal(r, g["r"])


def test_scheduler_tables(golden_dir):
    g = load(golden_dir, "scheduler_tables.pt")
    tab = S.Tables()
    for name, start, n in (("ddim50_s100", 100.0, 50), ("simple_orig100", 100.0, 100), ("ddim6_s20", 20.0, 6)):
        ts, sig, mvc = tab.ddim_schedule(start, None, n)
        assert torch.equal(ts, g[name]["timesteps"]), name
        assert torch.equal(sig, g[name]["sigmas"]), name
        assert torch.equal(tab.sigmas, g[name]["table"])
        assert float(mvc) == float(g[name]["min_var_coef"])


def test_denoise_loop_every_scheduler(golden_dir, tiny):
    sd, ssd = tiny
    g = load(golden_dir, "denoise_loop_tiny.pt")
    tab = S.Tables()
    fwd = lambda z, t: ddim_net.unet_forward(sd, z, t)
    enc = lambda z, t: ddim_net.unet_encode(sd, z, t)
    sgf = lambda f: ddim_net.sigma_forward(ssd, f)
    d = 3 * 16 * 16
    for key, case in g.items():
        kind, eta, var = key.split("|")
        ts, sig, mvc = tab.ddim_schedule(20.0, None, 6)
        assert torch.equal(ts, case["timesteps"])
        xT = case["z"] / (1 / (sig[0] ** 2 + 1)).sqrt()
        log = []
        with torch.no_grad():
            x0 = S.denoise_loop(tab, ts.tolist(), sig, mvc, fwd, enc, sgf, xT, kind=kind, eta=float(eta),
                                sampler_var=var, style="pred", norm_eps=True, refine=True, norm_min=0.0,
                                norm_max=30.0 / d ** 0.5, noises=case["noises"] or None, log=log)
        assert torch.equal(x0, case["final"]), key
        for i, st in enumerate(log):
            assert torch.equal(st["eps"], case["eps"][i]), (key, i)
            assert torch.equal(st["x0_hat"], case["x0_hat"][i]), (key, i)


def test_operators(golden_dir):
    g = load(golden_dir, "operators_r32.pt")
    R, C = 32, 3
    ops = {
        "inpainting": O.Inpainting(C, R, g["missing"]),
        "colorization": O.Colorization(R),
        "sr_averagepooling": O.SuperResolution(C, R, 4),
        "cs_walshhadamard": O.WalshHadamardCS(C, R, 4, g["perm"]),
        "sr_bicubic": O.SRConv(O.bicubic_kernel(4), C, R, 4),
        "deblur_gauss": O.Deblurring(O.gauss_kernel(), C, R),
    }
    for name, op in ops.items():
        y = op.A(g["x"].clone())
        assert torch.equal(y, g[name]["A"]), name
        assert torch.equal(op.At(y.clone()), g[name]["At"]), name
        assert torch.equal(op.A_pinv(y.clone()), g[name]["A_pinv"]), name
        assert torch.equal(op.project(g["x0"], y), g[name]["project"]), name


def test_operator_identities():
    """Known-answer properties of the reference operators: A A^+ y = y, projection feasibility and
    idempotence."""
    R, C, B = 32, 3, 2
    gen = torch.Generator().manual_seed(0)
    x = torch.rand(B, C * R * R, generator=gen) * 2 - 1
    x0 = torch.randn(B, C, R, R, generator=gen)
    perm = torch.randperm(R * R, generator=gen)
    for op, tol in ((O.Colorization(R), 5e-6), (O.SuperResolution(C, R, 4), 5e-6),
                    (O.WalshHadamardCS(C, R, 4, perm), 5e-6), (O.SRConv(O.bicubic_kernel(4), C, R, 4), 5e-5),
                    (O.Deblurring(O.gauss_kernel(), C, R), 5e-5)):
        y = op.A(x.clone())
        assert (op.A(op.A_pinv(y.clone())) - y).abs().max() < tol * 20
        p = op.project(x0, y)
        assert (op.A(p.reshape(B, -1)) - y).abs().max() < tol * 20
        assert (op.project(p, y) - p).abs().max() < 5e-4


@pytest.mark.parametrize("name", ["adm_tiny", "adm_alt"])
def test_adm_networks(golden_dir, name):
    from oracle import adm_net
    cfg = dict(weights.ADM_CONFIGS[name])
    sg = cfg.pop("sigma")
    sd = weights.adm_unet_state_dict(**cfg, seed=3)
    ssd = weights.adm_sigma_state_dict(**sg, seed=4)
    g = load(golden_dir, "nets_adm.pt")[name]
    with torch.no_grad():
        out, feat = adm_net.unet_forward(sd, g["x"], g["t"], cfg, return_feat=True)
        enc = adm_net.unet_encode(sd, g["x"], g["t"], cfg)
        r = adm_net.sigma_forward(ssd, feat, cfg)
    assert torch.equal(out, g["out"]) and torch.equal(feat, g["feat"]) and torch.equal(enc, g["feat"])
    assert torch.equal(r, g["r"])


def _edm_oracle():
    from oracle import edm_net, sampler_edm
    cfg = dict(weights.EDM_CONFIGS["edm_tiny"])
    sg = cfg.pop("sigma")
    sd = weights.edm_unet_state_dict(**cfg, seed=3)
    ssd = weights.edm_sigma_state_dict(**sg, seed=4)
    d = 3 * cfg["img_resolution"] ** 2
    o = sampler_edm.EDM(lambda x, c: edm_net.unet_forward(sd, x, c, cfg), lambda x, c: edm_net.unet_encode(sd, x, c, cfg),
                        lambda f: edm_net.sigma_forward(ssd, f), d, norm_min=0.0, norm_max=30.0 / d ** 0.5)
    return cfg, sd, ssd, o


def test_edm_networks(golden_dir):
    from oracle import edm_net
    cfg, sd, ssd, _ = _edm_oracle()
    g = load(golden_dir, "nets_edm.pt")
    with torch.no_grad():
        out, feat = edm_net.unet_forward(sd, g["x"], g["c_noise"], cfg, return_feat=True)
        r = edm_net.sigma_forward(ssd, feat)
    assert torch.equal(out, g["out"]) and torch.equal(feat, g["feat"]) and torch.equal(r, g["r"])


def test_edm_sampler_every_case(golden_dir):
    _, _, _, o = _edm_oracle()
    g = load(golden_dir, "edm_sampler_tiny.pt")
    for key, case in g.items():
        style, ne, refine, es = key.split("|")
        with torch.no_grad():
            x = o.sample(case["latents"], 4, style=style, norm_eps=ne + "0", refine=bool(int(refine)),
                         eps_scale=None if es == "None" else float(es))
        assert torch.equal(x, case["final"]), key


def test_continuous_t_dynamic_clip_and_projection_loop(golden_dir, tiny):
    """tests/golden/loops2_tiny.pt: the reference's denoise_loop with continuous_t / dynamic thresholding and the
    module-level projection_loop of image_sample.py (sigma feed-forward)."""
    sd, ssd = tiny
    g = load(golden_dir, "loops2_tiny.pt")
    fwd = lambda z, t: ddim_net.unet_forward(sd, z, t)
    enc = lambda z, t: ddim_net.unet_encode(sd, z, t)
    sgf = lambda f: ddim_net.sigma_forward(ssd, f)
    d = 3 * 16 * 16
    for key, case in g.items():
        loop, cont, clip, rates, kind, eta = key.split("|")
        tab = S.Tables()
        tab.continuous = bool(int(cont))
        ts, sig, mvc = tab.ddim_schedule(20.0, None, 6)
        assert torch.equal(ts.float(), case["timesteps"].float()) and torch.equal(sig, case["sigmas"].float())
        xT = case["z"] / (1 / (sig[0] ** 2 + 1)).sqrt()
        kw = dict(kind=kind, eta=float(eta), style="pred", norm_eps=True, refine=True, norm_min=0.0,
                  norm_max=30.0 / d ** 0.5, clip=clip, noises=case["noises"] or None)
        with torch.no_grad():
            if loop == "denoise":
                x0 = S.denoise_loop(tab, ts.tolist(), sig, mvc, fwd, enc, sgf, xT, **kw)
            else:
                x0 = S.projection_loop(tab, ts, sig, mvc, fwd, enc, sgf, xT, rates=eval(rates), **kw)
        assert torch.equal(x0, case["final"]), key


def test_anisotropic_deblurring(golden_dir):
    """tests/golden/operators2_r32.pt: functions/svd_operators.py Deblurring2D with the deblur_aniso kernels."""
    g = load(golden_dir, "operators2_r32.pt")
    op = O.Deblurring2D(*O.aniso_kernels(), 3, 32)
    ref = g["deblur_aniso"]
    y = op.A(g["x"])
    assert (y - ref["A"]).abs().max() < 5e-5
    assert (op.At(ref["A"]) - ref["At"]).abs().max() < 5e-5
    assert (op.A_pinv(ref["A"]) - ref["A_pinv"]).abs().max() < 5e-3 * ref["A_pinv"].abs().max()
    assert (op.project(g["x0"], ref["A"]) - ref["project"]).abs().max() < 5e-3 * ref["project"].abs().max()


def test_constrained_restoration_loops(golden_dir):
    """tests/golden/loops3_constrained.pt: the reference's denoise_loop with the DDNM `svd` projection and best-x0
    tracking (configs c4/c5: ADM learned-variance net, ddim_simple_orig, dynamic clip) for SR x4, box inpainting,
    colourisation and Walsh-Hadamard CS.  The operators run in spectral form here and the networks are the functional
    restatement, so the per-step tensors agree to fp32 round-off (1e-5), not bit for bit."""
    from functools import partial
    from oracle import adm_net
    g = load(golden_dir, "loops3_constrained.pt")
    cfg = dict(weights.ADM_CONFIGS["adm_tiny"])
    sg = cfg.pop("sigma")
    sd = weights.adm_unet_state_dict(**cfg, seed=3)
    ssd = weights.adm_sigma_state_dict(**sg, seed=4)
    R, C = cfg["image_size"], 3
    d = C * R * R
    fwd = lambda z, t: adm_net.unet_forward(sd, z, t, cfg)
    enc = lambda z, t: adm_net.unet_encode(sd, z, t, cfg)
    sgf = lambda f: adm_net.sigma_forward(ssd, f, cfg)
    for key, case in g.items():
        if "|" not in key:
            continue
        task, scale = key.split("|")
        op = {"sr_averagepooling": lambda: O.SuperResolution(C, R, int(scale)),
              "inpainting_box": lambda: O.Inpainting(C, R, g["missing"]),
              "colorization": lambda: O.Colorization(R),
              "cs_walshhadamard": lambda: O.WalshHadamardCS(C, R, int(scale), g["perm"])}[task]()
        deg = "inpainting" if task.startswith("inpainting") else task
        y = op.A(case["x_true"].reshape(2, -1))
        assert (y - case["y"].reshape(2, -1)).abs().max() < 1e-5, key
        tab = S.Tables()
        ts, sig, mvc = tab.ddim_schedule(20.0, None, 5)
        assert torch.equal(ts, case["timesteps"]) and torch.equal(sig, case["sigmas"].float())
        xT = case["z"] / (1 / (sig[0] ** 2 + 1)).sqrt()
        log = []
        with torch.no_grad():
            x0 = S.denoise_loop(tab, ts.tolist(), sig, mvc, fwd, enc, sgf, xT, kind="ddim_simple_orig", eta=0.85,
                                sampler_var="learned", style="pred", norm_eps=True, refine=True, norm_min=0.0,
                                norm_max=30.0 / d ** 0.5, clip="dynamic", noises=case["noises"],
                                constrain_fn=lambda v: op.project(v, case["y"]), sigma_pred_threshold=960,
                                learn_epsvar=True, log=log,
                                constrain_loss=lambda v: O.constraint_loss(op, deg, v, case["y"], C, R))
        for i, st in enumerate(log):
            for name in ("xt", "x0", "x_prev", "eps"):
                ref = case[name][i]
                assert (st[name] - ref).abs().max() <= 2e-5 * ref.abs().max(), (key, i, name)
            assert (st["const"] - case["const"][i]).abs().max() <= 1e-4 * case["const"][i].abs().max(), (key, i)
        assert (x0 - case["final"]).abs().max() <= 2e-5 * case["final"].abs().max(), key


def _ddnm_oracle_ops(g, R=32, C=3):
    return {"inpainting": O.Inpainting(C, R, g["missing"]), "colorization": O.Colorization(R),
            "sr_averagepooling": O.SuperResolution(C, R, 4), "cs_walshhadamard": O.WalshHadamardCS(C, R, 4, g["perm"]),
            "deblur_gauss": O.Deblurring(O.gauss_kernel(), C, R), "denoising": O.Denoising(C, R)}


def test_ddnm_plus_operator_terms(golden_dir):
    """tests/golden/ddnm_ops_r32.pt: Lambda / Lambda_noise / A_pinv_eta of the unmodified reference operators in five
    (a, sigma_y, sigma_t, eta) regimes.  The oracle's single spectral restatement reproduces every class bit for bit."""
    g = load(golden_dir, "ddnm_ops_r32.pt")
    for name, op in _ddnm_oracle_ops(g).items():
        for k, (a, sy, st, eta) in enumerate(g["regimes"]):
            a_t, st_t = torch.tensor(a), torch.tensor(st)
            assert torch.equal(op.Lambda(g["v"].clone(), a_t, sy, st_t, eta), g[name]["Lambda"][k]), (name, k)
            assert torch.equal(op.Lambda_noise(g["v"].clone(), a_t, sy, st_t, eta, g["e"].clone()),
                               g[name]["Lambda_noise"][k]), (name, k)
        if name != "denoising":
            for k, eta in enumerate((0.01, 0.5)):
                assert torch.equal(op.A_pinv_eta(g[name]["y"].clone(), eta), g[name]["A_pinv_eta"][k]), (name, k)
    for cls in (O.SRConv(O.bicubic_kernel(4), 3, 32, 4), O.Deblurring2D(*O.aniso_kernels(), 3, 32)):
        with pytest.raises(NotImplementedError):  # the reference defines no Lambda for these (base class raises)
            cls.Lambda(g["v"], 0.9, 0.1, 0.3, 0.85)


def test_ddnm_schedule_and_loops(golden_dir):
    """tests/golden/ddnm_loops_r32.pt: functions/svd_ddnm.py ddnm_diffusion / ddnm_plus_diffusion on the reference
    (adm_tiny network, T_sampling 4 with one time-travel detour).  Teacher-forced per step on the reference's own network
    outputs the oracle loop is exact to fp32 round-off; free-running through the functional network restatement 2e-5."""
    from oracle import adm_net, ddnm
    assert ddnm.schedule_jump(4, 2, 2) == [3, 2, 1, 0, 1, 2, 1, 0, -1]
    assert ddnm.schedule_jump(5, 1, 1) == [4, 3, 2, 1, 0, -1]
    assert ddnm.schedule_jump(10, 3, 3)[:12] == [9, 8, 7, 6, 7, 8, 9, 8, 7, 6, 7, 8]
    g = load(golden_dir, "ddnm_loops_r32.pt")
    cfg = dict(weights.ADM_CONFIGS["adm_tiny"])
    cfg.pop("sigma")
    sd = weights.adm_unet_state_dict(**cfg, seed=3)
    ops = _ddnm_oracle_ops(g)
    for key, case in g.items():
        if "|" not in key:
            continue
        name, sy = key.split("|")
        sy = None if sy == "None" else float(sy)
        zs = iter(case["z"])
        ets = iter(case["et"])
        rec = dict(xt=[], et=[], x0=[], x_next=[])
        # teacher-forced: the model returns the reference's recorded outputs
        x_last, x0_last = ddnm.run(case["xT"], lambda x, t: next(ets), g["betas"], g["eta"], ops[name], case["y"], sy,
                                   T_sampling=g["T_sampling"], travel_length=g["travel_length"],
                                   travel_repeat=g["travel_repeat"], noise_fn=lambda like: next(zs), record=rec)
        k = 0
        for i in range(len(rec["xt"])):
            if rec["et"][i].abs().max() == 0:
                continue  # time-travel step: no network call
            ref = case["xt"][k]
            assert (rec["xt"][i] - ref).abs().max() <= 2e-6 * ref.abs().max(), (key, i)
            k += 1
        assert k == len(case["xt"])
        assert (x_last - case["x_last"]).abs().max() <= 2e-6 * case["x_last"].abs().max(), key
        assert (x0_last - case["x0_last"]).abs().max() <= 2e-6 * case["x0_last"].abs().max(), key
        # free-running with the functional network
        zs = iter(case["z"])
        with torch.no_grad():
            x_last, x0_last = ddnm.run(case["xT"], lambda x, t: adm_net.unet_forward(sd, x, t, cfg), g["betas"], g["eta"],
                                       ops[name], case["y"], sy, T_sampling=g["T_sampling"],
                                       travel_length=g["travel_length"], travel_repeat=g["travel_repeat"],
                                       noise_fn=lambda like: next(zs))
        assert (x_last - case["x_last"]).abs().max() <= 2e-5 * case["x_last"].abs().max(), key


def test_block_cs_and_general_a(golden_dir):
    """tests/golden/operators3.pt: functions/svd_operators.py CS (with the reproducible Hadamard basis) and GeneralA."""
    g = load(golden_dir, "operators3.pt")
    V = O.hadamard_basis(1024, 7)
    assert torch.equal(V @ V.t(), torch.eye(1024))  # exactly orthogonal in fp32
    c = g["cs"]
    op = O.CS(3, 64, 0.25, V)
    assert torch.equal(op.A(c["x"]), c["A"]) and torch.equal(op.At(c["A"]), c["At"])
    assert torch.equal(op.A_pinv(c["A"]), c["A_pinv"]) and torch.equal(op.A_pinv_eta(c["A"], 0.1), c["A_pinv_eta"])
    assert torch.equal(op.project(c["x0"], c["A"]), c["project"])
    c = g["general"]
    op = O.GeneralA(c["Amat"].clone())
    for got, want in ((op.A(c["x"]), c["A"]), (op.At(c["A"]), c["At"]), (op.A_pinv(c["A"]), c["A_pinv"]),
                      (op.A_pinv_eta(c["A"], 0.1), c["A_pinv_eta"]), (op.project(c["x0"], c["A"]), c["project"])):
        assert (got - want).abs().max() <= 1e-5 * want.abs().max()  # the SVD is recomputed here


def test_ssim(golden_dir):
    """tests/golden/ssim.pt: image_sample.py:571-582 ssim_fn (uint8 rounding + basicsr's 3-D window SSIM) on the reference."""
    from oracle import metrics as M
    g = load(golden_dir, "ssim.pt")
    for size in g.values():
        for case in size.values():
            assert torch.equal(M.ssim3d(case["sample"], case["orig"]), case["ssim"])
    assert float(g["r32"]["same"]["ssim"].min()) == 1.0


def test_sigma_model_training_iteration(golden_dir):
    """tests/golden/train_step_tiny.pt: one training iteration of the reference's DDIM SigmaModel in train() mode
    (src/experiments.py:683-694): the oracle's functional forward with batch-statistics BatchNorm reproduces the loss, and
    autograd through it every parameter gradient, the AdamW update and the EMA copy (digests: L2 norm, sum, first entries;
    small tensors in full).  These are the targets a native backward pass has to hit."""
    from oracle import training as OT
    g = load(golden_dir, "train_step_tiny.pt")
    cfg = weights.CONFIGS["tiny"]
    ssd = weights.ddim_sigma_state_dict(**cfg["sigma"], seed=4)
    names = [n for n in g["grads"]]
    params = {n: torch.nn.Parameter(ssd[n].clone()) for n in names}
    sd = dict(ssd)
    sd.update(params)
    dist_hat = ddim_net.sigma_forward(sd, g["feat"], training=True) + 1
    assert (dist_hat.detach() - g["dist_hat"]).abs().max() <= 2e-6 * g["dist_hat"].abs().max()
    loss = torch.nn.functional.mse_loss(dist_hat, g["dist_real"])
    assert abs(loss.item() - g["loss"].item()) <= 2e-6 * abs(g["loss"].item())
    opt = OT.AdamWEma([params[n] for n in names], lr=1e-3, weight_decay=0.01, ema_rate=0.999)
    loss.backward()

    def check(t, ref, tol, what):
        t = t.detach().double().reshape(-1)
        scale = max(float(ref["norm"]), 1e-12)
        assert abs(float(t.norm()) - float(ref["norm"])) <= tol * scale, what
        assert (t[:32] - ref["head"]).abs().max() <= tol * max(float(ref["head"].abs().max()), scale / t.numel() ** 0.5), what
        if ref["full"] is not None:
            assert (t.float() - ref["full"]).abs().max() <= tol * max(float(ref["full"].abs().max()), 1e-12), what

    for n in names:
        check(params[n].grad, g["grads"][n], 2e-4, ("grad", n))
    opt.step()
    for n, e in zip(names, opt.ema):
        check(params[n], g["new_params"][n], 1e-5, ("param", n))
        check(e, g["ema"][n], 1e-5, ("ema", n))


# ------------------------------------------------------------------------------------------------------------------
# Benchmark architectures (tests/golden/make_golden.py bench_arch): the oracle at the REAL c2 / c3 / c4-c5 sizes
def _digest(t):
    v = t.double().reshape(-1)
    return torch.stack([v.sum(), (v * torch.arange(1, v.numel() + 1, dtype=torch.float64)).sum() / v.numel()])


def same_digest(t, want):
    """(float64 sums over ~1e5 values differ in the last bits between machines: compared to 1e-9 relative)"""
    return torch.allclose(_digest(t) if t.dim() != 1 or t.numel() != 2 else t, want, rtol=1e-9, atol=1e-9)


def bench_noise(shape, n, seed=5):
    """The reference's draws for a loop seeded with `seed` (x_T first, then one per step), regenerated instead of stored."""
    torch.manual_seed(seed)
    z = torch.randn(shape)
    return z, [torch.randn(shape) for _ in range(n)]


def test_c2_loop_snapshots_at_the_benchmark_architecture(golden_dir):
    """tests/golden/loop_c2_100.pt (config c2: CelebA-64 unet_ddim, 100 steps, batch 4): the regenerated noise matches the
    recorded digests, and the oracle's NLC step reproduces the reference's sigma_hat / eps / x_{t-1} bit for bit on the
    reference's own x_t at three of the stored steps (the whole 100-step loop is ~90 s of CPU and is left to the GPU)."""
    g = load(golden_dir, "loop_c2_100.pt")
    shape = (4, 3, 64, 64)
    z, noises = bench_noise(shape, 100)
    assert same_digest(z, g["z_digest"])
    assert all(same_digest(n, d) for n, d in zip(noises, g["noise_digest"]))
    cfg = weights.CONFIGS["c2"]
    sd = weights.ddim_unet_state_dict(**cfg["unet"], seed=3)
    ssd = weights.ddim_sigma_state_dict(**cfg["sigma"], seed=4)
    tab = S.Tables()
    ts, sig, mvc = tab.ddim_schedule(100.0, None, 100)
    assert torch.equal(ts, g["timesteps"]) and torch.equal(sig, g["sigmas"]) and torch.equal(tab.sigmas, g["table"])
    assert torch.equal(torch.searchsorted(tab.sigmas, g["sigma_t"].contiguous()), g["t_hat"])
    d = 3 * 64 * 64
    fwd = lambda z_, t: ddim_net.unet_forward(sd, z_, t)
    enc = lambda z_, t: ddim_net.unet_encode(sd, z_, t)
    sgf = lambda f: ddim_net.sigma_forward(ssd, f)
    for i in (0, 25, 99):
        sn = g["snap"][i]
        log = []
        with torch.no_grad():
            S.denoise_loop(tab, ts[i:i + 2].tolist(), sig[i:i + 2], mvc, fwd, enc, sgf, sn["xt"], kind="ddim_simple_orig",
                           eta=0.85, style="pred", norm_eps=True, refine=True, norm_min=-2.0 / d ** 0.5,
                           norm_max=110.0 / d ** 0.5, noises=[noises[i]], sigma_pred_threshold=960, log=log)
        st = log[0]
        assert torch.equal(st["sigma_t"], g["sigma_t"][i]) and torch.equal(st["sigma_prev"], g["sigma_prev"][i]), i
        assert torch.equal(st["eps"], sn["eps"]) and torch.equal(st["x0"], sn["x0"]), i
        assert torch.equal(st["x_prev"], sn["x_prev"]), i


def test_networks_at_the_benchmark_architectures(golden_dir):
    """tests/golden/nets_bench.pt: EDM SongUNet-64 (c3) and ADM-256 (c4/c5) outputs of the unmodified reference."""
    from oracle import adm_net, edm_net
    g = load(golden_dir, "nets_bench.pt")
    cfg = dict(weights.EDM_CONFIGS["edm64"])
    sg = cfg.pop("sigma")
    sd, ssd = weights.edm_unet_state_dict(**cfg, seed=3), weights.edm_sigma_state_dict(**sg, seed=4)
    e = g["edm64"]
    with torch.no_grad():
        out, feat = edm_net.unet_forward(sd, e["x"], e["c_noise"], cfg, return_feat=True)
        r = edm_net.sigma_forward(ssd, feat)
    assert torch.equal(out, e["out"]) and torch.equal(feat, e["feat"]) and torch.equal(r, e["r"])
    del sd, ssd
    cfg = dict(weights.ADM_CONFIGS["adm256"])
    sg = cfg.pop("sigma")
    sd, ssd = weights.adm_unet_state_dict(**cfg, seed=3), weights.adm_sigma_state_dict(**sg, seed=4)
    a = g["adm256"]
    torch.set_num_threads(8)
    with torch.no_grad():
        out, feat = adm_net.unet_forward(sd, a["x"], a["t"], cfg, return_feat=True)
        r = adm_net.sigma_forward(ssd, feat, cfg)
    torch.set_num_threads(4)
    assert torch.equal(out, a["out"]) and torch.equal(feat, a["feat"]) and torch.equal(r, a["r"])


@pytest.mark.parametrize("name", ["dhariwal_tiny", "dhariwal64"])
def test_dhariwal_unet(golden_dir, name):
    """tests/golden/nets_dhariwal.pt: DhariwalUNet.forward of the unmodified reference (src/edm_networks.py:406-502)."""
    from oracle import edm_net
    cfg = dict(weights.DHARIWAL_CONFIGS[name])
    sd = weights.dhariwal_unet_state_dict(**cfg, seed=3)
    g = load(golden_dir, "nets_dhariwal.pt")[name]
    with torch.no_grad():
        out = edm_net.dhariwal_forward(sd, g["x"], g["c_noise"], cfg)
    assert torch.equal(out, g["out"])
