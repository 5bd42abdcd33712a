"""Live pin of the oracle (and of the host-side scheduler mirror) against the unmodified reference imported from
/root/reference.  Skipped where the reference tree is absent (the GPU box); the committed golden vectors cover
the same ground there."""
import os

import numpy as np
import pytest
import torch

import refimport
from oracle import ddim_net, operators as O, sampler as S, weights

pytestmark = pytest.mark.skipif(not refimport.available(), reason="reference tree not present")
torch.set_num_threads(4)


@pytest.fixture(scope="module")
def R():
    return refimport.load()


def _make_golden():
    """tests/golden/make_golden.py as a module (it holds the builders of the reference modules)."""
    import importlib.util
    import os
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "make_golden.py")
    spec = importlib.util.spec_from_file_location("make_golden", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


@pytest.mark.parametrize("name", ["tiny", "c1"])
def test_state_dict_layout_and_networks(R, name):
    cfg = weights.CONFIGS[name]
    u, sg = cfg["unet"], cfg["sigma"]
    sd = weights.ddim_unet_state_dict(**u, seed=3)
    ssd = weights.ddim_sigma_state_dict(**sg, seed=4)
    net = R.unet_ddim.UNetModel(**u).eval()
    snet = R.unet_ddim.SigmaModel(dim=sg["dim"], channels=sg["channels"], n_blocks=sg["n_blocks"]).eval()
    for mine, ref in ((sd, net.state_dict()), (ssd, snet.state_dict())):
        assert set(mine) == set(ref)
        assert all(mine[k].shape == ref[k].shape for k in mine)
    net.load_state_dict(sd)
    snet.load_state_dict(ssd)
    x = torch.randn(2, 3, u["image_size"], u["image_size"])
    t = torch.tensor([640.0, 12.0])
    with torch.no_grad():
        assert torch.equal(net(x, t), ddim_net.unet_forward(sd, x, t))
        f = net.encode(x, t)
        assert torch.equal(f, ddim_net.unet_encode(sd, x, t))
        assert torch.equal(snet(f), ddim_net.sigma_forward(ssd, f))


def _simple_config(u, feat_layer):
    """The YAML-shaped config object src/unet_simple.py `Model(config)` reads (the store/config/*.yml files are absent)."""
    import types
    return types.SimpleNamespace(
        model=types.SimpleNamespace(ch=u["model_channels"], out_ch=u["out_channels"], ch_mult=list(u["channel_mult"]),
                                    num_res_blocks=u["num_res_blocks"], attn_resolutions=list(u["attention_resolutions"]),
                                    dropout=0.0, in_channels=u["in_channels"], resamp_with_conv=True, type="simple",
                                    feat_layer=feat_layer),
        data=types.SimpleNamespace(image_size=u["image_size"]),
        diffusion=types.SimpleNamespace(num_diffusion_timesteps=1000))


@pytest.mark.parametrize("feat_layer", [0, 1])
def test_unet_simple_feat_layer(R, feat_layer):
    """src/unet_simple.py `Model(config)` - the DDIM UNet the factory create_simple_sigma_eps_model actually builds
    (src/script_util.py:209-219) - has the state_dict of src/unet_ddim.py's UNetModel and a `feat_layer` switch: 0 returns
    the tensor after mid.attn_1, anything else the one after mid.block_2 (:371-375, :402-407)."""
    import importlib
    US = importlib.import_module("src.unet_simple")
    u = weights.CONFIGS["tiny"]["unet"]
    sd = weights.ddim_unet_state_dict(**u, seed=3)
    net = US.Model(_simple_config(u, feat_layer)).eval()
    assert set(net.state_dict()) == set(sd)
    net.load_state_dict(sd)
    x = torch.randn(2, 3, u["image_size"], u["image_size"], generator=torch.Generator().manual_seed(8))
    t = torch.tensor([640.0, 12.0])
    with torch.no_grad():
        assert torch.equal(net.encode(x, t), ddim_net.unet_encode(sd, x, t, feat_layer=feat_layer))
        out, feat = net.forward_and_encode(x, t)
        o2, f2 = ddim_net.unet_forward(sd, x, t, return_feat=True, feat_layer=feat_layer)
        assert torch.equal(out, o2) and torch.equal(feat, f2) and torch.equal(net(x, t), o2)


def test_c2_layout(R):
    cfg = weights.CONFIGS["c2"]
    sd = weights.ddim_unet_state_dict(**cfg["unet"], seed=3)
    ref = R.unet_ddim.UNetModel(**cfg["unet"]).state_dict()
    assert set(sd) == set(ref) and all(sd[k].shape == ref[k].shape for k in sd)


@pytest.mark.parametrize("kw", [
    dict(sampler_name="ddim", inference_timesteps=50, start_sigma=100),
    dict(sampler_name="ddim_simple_orig", inference_timesteps=100, eta=0.85),
    dict(sampler_name="ddim", inference_timesteps=10, sigma_style="EDM", start_sigma=80, end_sigma=0.02),
    dict(sampler_name="ddpm", inference_timesteps=20, sigma_style="Linear", start_sigma=50, end_sigma=0.05,
         sampler_var="fixedsmall"),
    dict(sampler_name="ddim", inference_timesteps=20, sigma_style="Scaled", start_sigma=50, end_sigma=0.05,
         linear_scale=1.1),
    dict(sampler_name="ddim", inference_timesteps=30, beta_schedule="cosine"),
    dict(sampler_name="ddim", inference_timesteps=50, start_sigma=100, continuous_t=True),
    dict(sampler_name="ddim", inference_timesteps=10, sigma_style="EDM", start_sigma=80, end_sigma=0.02, continuous_t=True),
    dict(sampler_name="ddpm", inference_timesteps=20, sigma_style="Linear", start_sigma=50, end_sigma=0.05,
         sampler_var="fixedsmall", continuous_t=True),
    dict(sampler_name="ddim", inference_timesteps=20, sigma_style="Linear", start_sigma=50.0, end_sigma=0.05),
    dict(sampler_name="ddim", inference_timesteps=20, sigma_style="Scaled", start_sigma=50, end_sigma=0.05,
         linear_scale=1.1, continuous_t=True),
])
def test_scheduler_mirror_tables(R, kw):
    from nlc_b200 import schedulers as M
    a = R.schedulers.get_sampler(train_timesteps=1000, **kw)
    b = M.get_sampler(train_timesteps=1000, **kw)
    assert a.timesteps.dtype == b.timesteps.dtype and torch.equal(a.timesteps, b.timesteps)
    assert a.sampling_sigmas.dtype == b.sampling_sigmas.dtype and torch.equal(a.sampling_sigmas, b.sampling_sigmas)
    s = torch.rand(9) * 90 + 0.01
    assert torch.equal(a.get_t_from_sigma(s.view(-1, 1, 1, 1)).reshape(-1), b.get_t_from_sigma(s.view(-1, 1, 1, 1)).reshape(-1))
    assert torch.equal(a.sigmas, b.sigmas)
    assert float(a.min_var_coef) == float(b.min_var_coef)


@pytest.mark.parametrize("kind,eta,var", [("ddim", 0.0, "none"), ("ddim_simple_orig", 0.85, "none"),
                                          ("ddpm", 1.0, "fixedlarge"), ("ddpm_orig", 1.0, "fixedsmall")])
def test_denoise_loop(R, kind, eta, var):
    cfg = weights.CONFIGS["tiny"]
    u, sg = cfg["unet"], cfg["sigma"]
    sd = weights.ddim_unet_state_dict(**u, seed=3)
    ssd = weights.ddim_sigma_state_dict(**sg, seed=4)
    net = R.unet_ddim.UNetModel(**u).eval()
    net.load_state_dict(sd)
    snet = R.unet_ddim.SigmaModel(dim=sg["dim"], channels=sg["channels"], n_blocks=sg["n_blocks"]).eval()
    snet.load_state_dict(ssd)
    sch = R.schedulers.get_sampler(kind, 1000, 5, start_sigma=30.0, sampler_var=var, eta=eta)
    B, shape = 2, (2, 3, 16, 16)
    exp = R.experiments.ImageExperiment(net, sch, batch_size=B, data_shape=shape[1:], seed=9, device="cpu")
    exp.set_model(net, snet, learn_epsvar=False)
    exp.set_norm_maxmin(-2.0, 25.0)
    exp.set_clip_fn("clamp")
    out, _ = exp.denoise_loop(shape=shape, gen=exp.new_gen(9), style="pred", norm_eps=True, refine_prior_sigma=True,
                              return_log=False, chunk_size=1)
    tab = S.Tables()
    ts, sig, mvc = tab.ddim_schedule(30.0, None, 5)
    torch.manual_seed(9)
    z = torch.randn(shape)
    noises = [torch.randn(shape) for _ in range(len(ts) - 1)] if (eta > 0 or kind.startswith("ddpm")) else None
    with torch.no_grad():
        x0 = S.denoise_loop(tab, ts.tolist(), sig, mvc, lambda a, t: ddim_net.unet_forward(sd, a, t),
                            lambda a, t: ddim_net.unet_encode(sd, a, t), lambda f: ddim_net.sigma_forward(ssd, f),
                            z / (1 / (sig[0] ** 2 + 1)).sqrt(), kind=kind, eta=eta, sampler_var=var, style="pred",
                            norm_eps=True, refine=True, norm_min=exp.norm_min, norm_max=exp.norm_max, noises=noises)
    assert torch.equal(x0, out)


def test_operators(R):
    ref = R.svd_operators
    Rr, C, B = 32, 3, 2
    x = torch.rand(B, C * Rr * Rr) * 2 - 1
    x0 = torch.randn(B, C, Rr, Rr)
    mask = torch.ones(Rr, Rr)
    mask[4:20, 10:30] = 0
    mr = torch.nonzero(mask.reshape(-1) == 0).long().reshape(-1) * 3
    missing = torch.cat([mr, mr + 1, mr + 2])
    perm = torch.randperm(Rr * Rr)
    pairs = [
        (ref.Inpainting(C, Rr, missing, "cpu"), O.Inpainting(C, Rr, missing)),
        (ref.Colorization(Rr, "cpu"), O.Colorization(Rr)),
        (ref.SuperResolution(C, Rr, 2, "cpu"), O.SuperResolution(C, Rr, 2)),
        (ref.WalshHadamardCS(C, Rr, 2, perm, "cpu"), O.WalshHadamardCS(C, Rr, 2, perm)),
        (ref.SRConv(O.bicubic_kernel(2), C, Rr, "cpu", stride=2), O.SRConv(O.bicubic_kernel(2), C, Rr, 2)),
        (ref.Deblurring(O.gauss_kernel(), C, Rr, "cpu"), O.Deblurring(O.gauss_kernel(), C, Rr)),
        (ref.Deblurring(torch.Tensor([1 / 9] * 9), C, Rr, "cpu"), O.Deblurring(torch.Tensor([1 / 9] * 9), C, Rr)),
        (ref.Deblurring2D(*O.aniso_kernels(), C, Rr, "cpu"), O.Deblurring2D(*O.aniso_kernels(), C, Rr)),
    ]
    for a, b in pairs:
        y = a.A(x.clone())
        assert torch.equal(y, b.A(x.clone()))
        assert torch.equal(a.At(y.clone()), b.At(y.clone()))
        assert torch.equal(a.A_pinv(y.clone()), b.A_pinv(y.clone()))
        pr = x0 - a.A_pinv(a.A(x0.reshape(B, -1)) - y).reshape(x0.shape)
        assert torch.equal(pr, b.project(x0, y))
        assert torch.equal(a.A_pinv_eta(y.clone(), 0.05), b.A_pinv_eta(y.clone(), 0.05))
        if not hasattr(a, "_singulars_orig") and type(a).__name__ in ("SRConv", "Deblurring2D"):
            continue  # no Lambda in the reference
        # DDNM+ terms on random regimes: 0-dim fp32 a / sigma_t as functions/svd_ddnm.py passes them
        v, e = torch.randn(B, C * Rr * Rr), torch.randn(B, C * Rr * Rr)
        for _ in range(4):
            a_t, st_t = torch.rand(()) * 0.9 + 0.1, torch.rand(()) * 0.9 + 0.05
            sy, eta = float(torch.rand(())) * 0.5, float(torch.rand(()))
            assert torch.equal(a.Lambda(v.clone(), a_t, sy, st_t, eta), b.Lambda(v.clone(), a_t, sy, st_t, eta))
            assert torch.equal(a.Lambda_noise(v.clone(), a_t, sy, st_t, eta, e.clone()),
                               b.Lambda_noise(v.clone(), a_t, sy, st_t, eta, e.clone()))


def test_block_cs_general_a_denoising(R):
    """The remaining operator classes of functions/svd_operators.py (:101-208, 442-476), live, with the reference's own
    random basis."""
    ref = R.svd_operators
    B = 2
    a = ref.CS(3, 64, 0.25, "cpu")
    pairs = [(a, O.CS(3, 64, 0.25, a.V_small), 3 * 64 * 64)]
    Am = torch.randn(24, 80)
    pairs.append((ref.GeneralA(Am.clone()), O.GeneralA(Am.clone()), 80))
    pairs.append((ref.Denoising(3, 32, "cpu"), O.Denoising(3, 32), 3 * 32 * 32))
    for a, b, n in pairs:
        x, x0 = torch.rand(B, n) * 2 - 1, torch.randn(B, n)
        y = a.A(x.clone())
        assert torch.equal(y, b.A(x.clone()))
        assert torch.equal(a.At(y.clone()), b.At(y.clone()))
        assert torch.equal(a.A_pinv(y.clone()), b.A_pinv(y.clone()))
        assert torch.equal(a.A_pinv_eta(y.clone(), 0.3), b.A_pinv_eta(y.clone(), 0.3))
        assert torch.equal(x0 - a.A_pinv(a.A(x0.clone()) - y), b.project(x0, y))


def test_training_batch_preparation(R):
    """oracle/training.prepare_batch against the reference's own lines (src/experiments.py:666-669 with its
    Scheduler.diffusion and vector_norm), live."""
    from oracle import training as OT
    sch = R.schedulers.get_sampler("ddim", 1000, 10)
    B, shape = 4, (4, 3, 16, 16)
    x0, noise, extra = torch.rand(shape) * 2 - 1, torch.randn(shape), torch.randn(shape)
    eta1, eta2 = 0.05 + torch.rand(B, 1, 1, 1) * 0.2, 0.1 + torch.rand(B, 1, 1, 1)
    t = torch.randint(0, 1000, (B,))
    noise_delta = eta1 * noise + eta1 * eta2 * extra
    new_noise = noise + noise_delta
    dist_real = R.utils.vector_norm(new_noise) / np.sqrt(3 * 16 * 16)
    noisy_x, _ = sch.diffusion(x0, t, new_noise)
    got_x, got_d, got_n = OT.prepare_batch(x0, t, noise, extra, eta1, eta2, sch.alphas_cumprod)
    assert torch.equal(got_x, noisy_x) and torch.equal(got_d, dist_real) and torch.equal(got_n, new_noise)


def test_operator_edge_cases(R):
    """Degenerate instances the reference accepts: nothing / everything missing, ratio 1, a rank-deficient general matrix
    (zero-guarded pseudo-inverse), batch 1."""
    ref = R.svd_operators
    Rr, C = 32, 3
    n = C * Rr * Rr
    none, every = torch.zeros(0, dtype=torch.long), torch.arange(n)
    Am = torch.randn(10, 40)
    Am[7] = Am[2]  # rank 9: one singular value falls under the 1e-3 threshold
    pairs = [
        (ref.Inpainting(C, Rr, none, "cpu"), O.Inpainting(C, Rr, none), n),
        (ref.Inpainting(C, Rr, every, "cpu"), O.Inpainting(C, Rr, every), n),
        (ref.WalshHadamardCS(C, Rr, 1, torch.randperm(Rr * Rr), "cpu"), None, n),
        (ref.SuperResolution(C, Rr, 1, "cpu"), O.SuperResolution(C, Rr, 1), n),
        (ref.GeneralA(Am.clone()), O.GeneralA(Am.clone()), 40),
    ]
    pairs[2] = (pairs[2][0], O.WalshHadamardCS(C, Rr, 1, pairs[2][0].perm), n)
    for B in (1, 3):
        for a, b, d in pairs:
            x, x0 = torch.rand(B, d) * 2 - 1, torch.randn(B, d)
            y = a.A(x.clone())
            assert y.shape == b.A(x.clone()).shape and torch.equal(y, b.A(x.clone()))
            assert torch.equal(a.At(y.clone()), b.At(y.clone()))
            assert torch.equal(a.A_pinv(y.clone()), b.A_pinv(y.clone()))
            assert torch.equal(x0 - a.A_pinv(a.A(x0.clone()) - y), b.project(x0, y))
    assert pairs[1][0].A(torch.zeros(2, n)).shape == (2, 0)  # everything missing: an empty measurement


def test_ssim(R):
    """oracle/metrics.ssim3d against the unmodified basicsr code behind image_sample.py:571-582, live, random sizes."""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
    import make_golden as MG
    from oracle import metrics as M
    PS = MG.basicsr_psnr_ssim()
    for H, W in ((32, 32), (17, 45), (64, 64)):
        a, b = torch.rand(2, 3, H, W), torch.rand(2, 3, H, W)
        b = 0.7 * a + 0.3 * b
        assert M.ssim3d(a, b).tolist() == MG.reference_ssim_fn(PS, a, b)


@pytest.mark.parametrize("name", ["adm_tiny", "adm_alt"])
def test_adm_networks(R, name):
    """oracle/adm_net.py == src/unet_adm.py (UNetModel forward/encode, SigmaModel), bit for bit."""
    from oracle import adm_net
    cfg, sg, sd, ssd, net, snet = _make_golden().adm_reference_modules(name)
    for mine, ref in ((sd, net.state_dict()), (ssd, snet.state_dict())):
        assert set(mine) == set(ref)
        assert all(mine[k].shape == ref[k].shape for k in mine)
    x = torch.randn(2, 3, cfg["image_size"], cfg["image_size"])
    t = torch.tensor([640.0, 12.0])
    with torch.no_grad():
        assert torch.equal(net(x, t), adm_net.unet_forward(sd, x, t, cfg))
        f = net.encode(x, t)
        assert torch.equal(f, adm_net.unet_encode(sd, x, t, cfg))
        assert torch.equal(snet(f), adm_net.sigma_forward(ssd, f, cfg))


@pytest.mark.parametrize("name", ["adm_tiny", "adm_alt"])
def test_adm_sigma_model_train_mode_and_gradients(R, name):
    """The oracle's TRAIN-mode ADM sigma-model forward (batch-statistics BatchNorm1d; dropout 0) and its autograd gradients
    against the reference's own module in train mode: the target nlc_b200.training.NativeSigmaModel(family="adm") is held to
    on the GPU (tests/test_gpu_training.py)."""
    from oracle import adm_net
    cfg, sg, sd, ssd, net, snet = _make_golden().adm_reference_modules(name)
    g = torch.Generator().manual_seed(6)
    feat = torch.randn(5, sg["channels"], sg["dim"], sg["dim"], generator=g)
    target = 1.0 + 0.3 * torch.randn(5, 1, 1, 1, generator=g)
    snet.train()
    for mod in snet.modules():
        if isinstance(mod, torch.nn.Dropout):
            mod.p = 0.0
    ref = torch.nn.functional.mse_loss(snet(feat) + 1, target)
    ref.backward()
    names = [k for k, _ in snet.named_parameters()]
    params = {n: torch.nn.Parameter(ssd[n].clone()) for n in names}
    mine = dict(ssd)
    mine.update(params)
    loss = torch.nn.functional.mse_loss(adm_net.sigma_forward(mine, feat, cfg, training=True) + 1, target)
    loss.backward()
    assert torch.allclose(loss, ref, rtol=1e-6, atol=0)
    gmax = max(float(p.grad.norm()) for p in snet.parameters())
    for n, p_ref in snet.named_parameters():
        assert (params[n].grad - p_ref.grad).norm() <= 1e-5 * p_ref.grad.norm() + 1e-7 * gmax, n


def test_edm_sigma_model_train_mode_and_gradients(R):
    """The oracle's train-mode EDM sigma-model forward and its autograd gradients against the reference's own module in train
    mode (dropout 0): the target of NativeSigmaModel(family="edm") on the GPU.  (`norm1` of a PureUNetBlock is never applied,
    src/edm_networks.py:940-944: its parameters get no gradient on either side.)"""
    from oracle import edm_net
    cfg, sg, sd, ssd, net, snet = _make_golden().edm_reference_modules("edm_tiny")
    g = torch.Generator().manual_seed(6)
    feat = torch.randn(5, sg["channels"], sg["dim"], sg["dim"], generator=g)
    target = 1.0 + 0.3 * torch.randn(5, 1, 1, 1, generator=g)
    snet.train()
    for mod in snet.modules():
        if hasattr(mod, "dropout") and isinstance(mod.dropout, float):
            mod.dropout = 0.0
    ref = torch.nn.functional.mse_loss(snet(feat.clone()) + 1, target)
    ref.backward()
    names = [k for k, _ in snet.named_parameters()]
    params = {n: torch.nn.Parameter(ssd[n].clone()) for n in names}
    mine = dict(ssd)
    mine.update(params)
    loss = torch.nn.functional.mse_loss(edm_net.sigma_forward(mine, feat.clone(), training=True) + 1, target)
    loss.backward()
    assert torch.allclose(loss, ref, rtol=1e-6, atol=0)
    used = [(n, p) for n, p in snet.named_parameters() if p.grad is not None]
    assert {n for n, _ in snet.named_parameters() if n not in dict(used)} == {n for n in names if params[n].grad is None}
    gmax = max(float(p.grad.norm()) for _, p in used)
    for n, p_ref in used:
        assert (params[n].grad - p_ref.grad).norm() <= 1e-5 * p_ref.grad.norm() + 1e-7 * gmax, n


def test_adm256_layout(R):
    """Key names / shapes of the c4/c5 architecture (no forward: 553 M parameters)."""
    import importlib
    UA = importlib.import_module("src.unet_adm")
    cfg = dict(weights.ADM_CONFIGS["adm256"])
    cfg.pop("sigma")
    keys = ("image_size", "model_channels", "out_channels", "num_res_blocks", "attention_resolutions", "channel_mult",
            "num_heads", "num_head_channels", "use_scale_shift_norm", "resblock_updown", "use_new_attention_order")
    with torch.device("meta"):
        ref = UA.UNetModel(in_channels=3, **{k: cfg[k] for k in keys}).state_dict()
    shapes = weights.adm_unet_state_dict(**cfg, seed=0, shapes_only=True)
    assert set(shapes) == set(ref) and all(tuple(shapes[k]) == tuple(ref[k].shape) for k in shapes)


def test_edm_networks(R):
    """oracle/edm_net.py == src/edm_networks.py (SongUNet forward/encode, SigmaModel), bit for bit."""
    from oracle import edm_net
    cfg, sg, sd, ssd, net, snet = _make_golden().edm_reference_modules("edm_tiny")
    for mine, ref in ((sd, net.state_dict()), (ssd, snet.state_dict())):
        assert set(mine) == set(ref)
        assert all(mine[k].shape == ref[k].shape for k in mine)
    x = torch.randn(2, 3, cfg["img_resolution"], cfg["img_resolution"])
    c = torch.tensor([0.9, -1.2])
    with torch.no_grad():
        assert torch.equal(net(x, c, None), edm_net.unet_forward(sd, x, c, cfg))
        f = net.encode(x, c, None)
        assert torch.equal(f, edm_net.unet_encode(sd, x, c, cfg))
        assert torch.equal(snet(f), edm_net.sigma_forward(ssd, f))


def test_edm64_layout(R):
    import importlib
    EN = importlib.import_module("src.edm_networks")
    cfg = dict(weights.EDM_CONFIGS["edm64"])
    sg = cfg.pop("sigma")
    ref = EN.SongUNet(**{k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.items()}).state_dict()
    sd = weights.edm_unet_state_dict(**cfg, seed=0)
    assert set(sd) == set(ref) and all(sd[k].shape == ref[k].shape for k in sd)
    sref = EN.SigmaModel(dim=sg["dim"], channels=sg["channels"], n_blocks=sg["n_blocks"]).state_dict()
    ssd = weights.edm_sigma_state_dict(**sg, seed=0)
    assert set(ssd) == set(sref) and all(ssd[k].shape == sref[k].shape for k in ssd)


@pytest.mark.parametrize("style,ne,refine,es", [("pred_partial,pred", "00", False, 1.0), ("pred,pred_partial", "11", True, 1.0),
                                                ("pred_sigma,pred_partial3", "10", False, None)])
def test_edm_sampler(R, style, ne, refine, es):
    """oracle/sampler_edm.py == EDMImageExperiment.edm_sampler (float64 result, torch.equal)."""
    from oracle import edm_net, sampler_edm
    cfg, sg, sd, ssd, net, snet = _make_golden().edm_reference_modules("edm_tiny")
    B, Rr = 2, cfg["img_resolution"]
    exp = R.experiments.EDMImageExperiment(net, None, batch_size=B, data_shape=(3, Rr, Rr), seed=1, device="cpu",
                                           num_timesteps=3, sigma_min=0.002, sigma_max=80)
    exp.set_model(net, snet, learn_epsvar=False)
    exp.set_norm_maxmin(0.0, 30.0)
    o = sampler_edm.EDM(lambda x, c: edm_net.unet_forward(sd, x, c, cfg), lambda x, c: edm_net.unet_encode(sd, x, c, cfg),
                        lambda f: edm_net.sigma_forward(ssd, f), 3 * Rr * Rr, norm_min=exp.norm_min, norm_max=exp.norm_max)
    with torch.no_grad():
        ref = exp.edm_sampler((B, 3, Rr, Rr), gen=R.experiments.StackedRandomGenerator("cpu", [5, 6]), style=style,
                              norm_eps=ne + "0", refine_prior_sigma=refine, eps_scale=es)
        lat = R.experiments.StackedRandomGenerator("cpu", [5, 6]).randn((B, 3, Rr, Rr), device="cpu")
        mine = o.sample(lat, 3, style=style, norm_eps=ne + "0", refine=refine, eps_scale=es)
    assert ref.dtype == torch.float64 and torch.equal(ref, mine)


@pytest.mark.parametrize("cont,clip,rates,kind,eta", [(False, "clamp", [1, 0, 0, 0], "ddim", 0.0),
                                                      (True, "dynamic", [0.5, 0.2, 0.2, 0.1], "ddim_simple_orig", 0.85)])
def test_projection_loop_and_continuous_t(R, cont, clip, rates, kind, eta):
    """oracle/sampler.py projection_loop / continuous-t denoise_loop == image_sample.projection_loop /
    ImageExperiment.denoise_loop of the reference (torch.equal)."""
    IS = _make_golden().image_sample_module()
    cfg = weights.CONFIGS["tiny"]
    u, sg = cfg["unet"], cfg["sigma"]
    sd = weights.ddim_unet_state_dict(**u, seed=3)
    ssd = weights.ddim_sigma_state_dict(**sg, seed=4)
    net = R.unet_ddim.UNetModel(**u).eval()
    net.load_state_dict(sd)
    snet = R.unet_ddim.SigmaModel(dim=sg["dim"], channels=sg["channels"], n_blocks=sg["n_blocks"]).eval()
    snet.load_state_dict(ssd)
    shape = (2, 3, 16, 16)
    sch = R.schedulers.get_sampler(kind, 1000, 4, start_sigma=30.0, eta=eta, continuous_t=cont)
    exp = R.experiments.ImageExperiment(net, sch, batch_size=2, data_shape=shape[1:], seed=9, device="cpu")
    exp.set_model(net, snet, learn_epsvar=False)
    exp.set_norm_maxmin(-2.0, 25.0)
    exp.set_clip_fn(clip)
    kw = dict(shape=shape, style="pred", norm_eps=True, refine_prior_sigma=True, chunk_size=1)
    ref_p, _ = IS.projection_loop(exp, gen=exp.new_gen(9), sigma_estimate_rate=rates, **kw)
    ref_d, _ = exp.denoise_loop(gen=exp.new_gen(9), return_log=False, **kw)
    tab = S.Tables()
    tab.continuous = cont
    ts, sig, mvc = tab.ddim_schedule(30.0, None, 4)
    torch.manual_seed(9)
    z = torch.randn(shape)
    noises = [torch.randn(shape) for _ in range(len(ts) - 1)] if eta > 0 else None
    f = (lambda a, t: ddim_net.unet_forward(sd, a, t), lambda a, t: ddim_net.unet_encode(sd, a, t),
         lambda x: ddim_net.sigma_forward(ssd, x))
    okw = dict(kind=kind, eta=eta, style="pred", norm_eps=True, refine=True, norm_min=exp.norm_min,
               norm_max=exp.norm_max, clip=clip, noises=noises)
    xT = z / (1 / (sig[0] ** 2 + 1)).sqrt()
    with torch.no_grad():
        assert torch.equal(S.projection_loop(tab, ts, sig, mvc, *f, xT, rates=rates, **okw), ref_p)
        assert torch.equal(S.denoise_loop(tab, ts.tolist(), sig, mvc, *f, xT, **okw), ref_d)
