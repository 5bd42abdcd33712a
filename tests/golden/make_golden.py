"""Generate the golden fixtures in this directory from the UNMODIFIED reference (imported read-only from
/root/reference).  Run here (the reference does not exist on the GPU box):

    python tests/golden/make_golden.py

Fixtures are small float32 tensors; weights are not stored — they are regenerated from oracle/weights.py seeds
(whose key names and shapes are checked against the reference modules by tests/test_oracle_vs_reference.py).
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import refimport  # noqa: E402
from oracle import operators as O  # noqa: E402  (only for the blur kernels' construction helpers)
from oracle import weights  # noqa: E402

SAMPLER_CASES = [("ddim", 0.0, "none"), ("ddim_simple_orig", 0.85, "none"), ("ddim", 0.5, "fixedsmall"),
                 ("ddpm", 1.0, "fixedlarge"), ("ddpm_orig", 1.0, "fixedsmall"), ("ddim_orig", 0.3, "fixedlarge"),
                 ("ddim_simple", 0.2, "none"), ("ddim_simple_drag", 0.2, "none")]


def main():
    R = refimport.load()
    torch.set_num_threads(4)
    cfg = weights.CONFIGS["tiny"]
    u, sg = cfg["unet"], cfg["sigma"]
    sd = weights.ddim_unet_state_dict(**u, seed=3)
    ssd = weights.ddim_sigma_state_dict(**sg, seed=4)
    net = R.unet_ddim.UNetModel(**u).eval()
    net.load_state_dict(sd)
    snet = R.unet_ddim.SigmaModel(dim=sg["dim"], channels=sg["channels"], n_blocks=sg["n_blocks"]).eval()
    snet.load_state_dict(ssd)

    # ---- networks
    g = torch.Generator().manual_seed(11)
    x = torch.randn(2, 3, u["image_size"], u["image_size"], generator=g)
    t = torch.tensor([500.0, 37.0])
    with torch.no_grad():
        out = net(x, t)
        feat = net.encode(x, t)
        r = snet(feat)
    torch.save(dict(x=x, t=t, out=out, feat=feat, r=r), os.path.join(HERE, "nets_tiny.pt"))

    # ---- scheduler tables
    tabs = {}
    for name, kw in (("ddim50_s100", dict(sampler_name="ddim", inference_timesteps=50, start_sigma=100)),
                     ("simple_orig100", dict(sampler_name="ddim_simple_orig", inference_timesteps=100, start_sigma=100,
                                             eta=0.85)),
                     ("ddim6_s20", dict(sampler_name="ddim", inference_timesteps=6, start_sigma=20.0))):
        s = R.schedulers.get_sampler(train_timesteps=1000, **kw)
        tabs[name] = dict(timesteps=s.timesteps.clone(), sigmas=s.sampling_sigmas.clone().float(),
                          min_var_coef=torch.as_tensor(s.min_var_coef).clone().float(), table=s.sigmas.clone())
    torch.save(tabs, os.path.join(HERE, "scheduler_tables.pt"))

    # ---- the reference's own denoise_loop, per-step dumps
    B, side = 2, u["image_size"]
    shape = (B, 3, side, side)
    loops = {}
    for kind, eta, var in SAMPLER_CASES:
        sch = R.schedulers.get_sampler(kind, 1000, 6, start_sigma=20.0, sampler_var=var, eta=eta)
        sch.to("cpu")
        exp = R.experiments.ImageExperiment(net, sch, batch_size=B, data_shape=(3, side, side), seed=5, device="cpu")
        exp.set_model(net, snet, learn_epsvar=False)
        exp.set_norm_maxmin(0.0, 30.0)
        exp.set_clip_fn("clamp")
        # the reference reseeds the *default* generator (new_gen -> torch.manual_seed) and then draws x_T and every
        # step's randn_like from it; record the draws by replaying the same stream
        torch.manual_seed(5)
        z = torch.randn(shape)
        need = kind in ("ddpm", "ddpm_orig") or eta > 0
        noises = [torch.randn(shape) for _ in range(len(sch.timesteps) - 1)] if need else []
        # observe (not modify) the reference's own update call: its inputs carry x_t and the corrected sigmas
        rec = dict(xt=[], sigma_t=[], sigma_prev=[], x_prev=[])
        orig = sch.pred_xprev

        def spy(*a, _orig=orig, _rec=rec, **k):
            out = _orig(*a, **k)
            _rec["xt"].append(k["xt"].clone())
            _rec["sigma_t"].append(torch.as_tensor(k["sigma_t"]).reshape(-1).clone())
            _rec["sigma_prev"].append(torch.as_tensor(k["sigma_prev"]).reshape(-1).clone())
            _rec["x_prev"].append(out.clone())
            return out

        sch.pred_xprev = spy
        # the two discrete time lookups of every step, observed (src/experiments.py:410,427)
        t_first, t_second = [], []
        orig_enc, orig_pred = exp.encode_xt, exp.pred_xt

        def enc_spy(xt, t, sigma=None, batch_t=True, _o=orig_enc, _l=t_first):
            _l.append(torch.as_tensor(t).reshape(-1).float().expand(B).clone())
            return _o(xt, t, sigma=sigma, batch_t=batch_t)

        def pred_spy(xt, t, sigma=None, batch_t=True, _o=orig_pred, _l=t_second):
            _l.append(torch.as_tensor(t).reshape(-1).float().expand(B).clone())
            return _o(xt, t, sigma=sigma, batch_t=batch_t)

        exp.encode_xt, exp.pred_xt = enc_spy, pred_spy
        final, logs = exp.denoise_loop(shape=shape, gen=exp.new_gen(5), style="pred", norm_eps=True,
                                       refine_prior_sigma=True, return_log=True, chunk_size=1)
        loops["%s|%s|%s" % (kind, eta, var)] = dict(
            t_first=torch.stack(t_first), t_hat=torch.stack(t_second),
            z=z, noises=noises, final=final, eps=logs[1], x0_hat=logs[2], x0=logs[3],
            xt=rec["xt"], sigma_t=rec["sigma_t"], sigma_prev=rec["sigma_prev"], x_prev=rec["x_prev"],
            timesteps=sch.timesteps.clone(), sigmas=sch.sampling_sigmas.clone())
    torch.save(loops, os.path.join(HERE, "denoise_loop_tiny.pt"))
    if os.environ.get("NLC_GOLDEN_STOP_AFTER_LOOPS") == "1":  # (regenerate the first three fixtures only)
        return

    # ---- operators
    ref = R.svd_operators
    Rr, C, Bo = 32, 3, 2
    g = torch.Generator().manual_seed(21)
    xs = torch.rand(Bo, C * Rr * Rr, generator=g) * 2 - 1
    x0 = torch.randn(Bo, C, Rr, Rr, generator=g)
    mask = torch.ones(Rr, Rr)
    mask[8:24, 8:24] = 0
    mr = torch.nonzero(mask.reshape(-1) == 0).long().reshape(-1) * 3
    missing = torch.cat([mr, mr + 1, mr + 2])
    perm = torch.randperm(Rr * Rr, generator=torch.Generator().manual_seed(3))
    ops = {
        "inpainting": ref.Inpainting(C, Rr, missing, "cpu"),
        "colorization": ref.Colorization(Rr, "cpu"),
        "sr_averagepooling": ref.SuperResolution(C, Rr, 4, "cpu"),
        "cs_walshhadamard": ref.WalshHadamardCS(C, Rr, 4, perm, "cpu"),
        "sr_bicubic": ref.SRConv(O.bicubic_kernel(4), C, Rr, "cpu", stride=4),
        "deblur_gauss": ref.Deblurring(O.gauss_kernel(), C, Rr, "cpu"),
    }
    gold = dict(x=xs, x0=x0, missing=missing, perm=perm)
    for name, op in ops.items():
        y = op.A(xs.clone())
        proj = x0 - op.A_pinv(op.A(x0.reshape(Bo, -1)) - y).reshape(x0.shape)
        gold[name] = dict(A=y, At=op.At(y.clone()), A_pinv=op.A_pinv(y.clone()), project=proj)
    torch.save(gold, os.path.join(HERE, "operators_r32.pt"))
    adm()
    edm()
    loops2()
    loops3()
    operators2()
    for f in sorted(os.listdir(HERE)):
        if f.endswith(".pt"):
            print(f, os.path.getsize(os.path.join(HERE, f)))


def adm_reference_modules(name):
    """The reference's ADM UNet + sigma-model for oracle/weights.ADM_CONFIGS[name], loaded with the seeded weights."""
    import importlib
    refimport.load()
    UA = importlib.import_module("src.unet_adm")
    cfg = dict(weights.ADM_CONFIGS[name])
    sg = cfg.pop("sigma")
    sd = weights.adm_unet_state_dict(**cfg, seed=3)
    ssd = weights.adm_sigma_state_dict(**sg, seed=4)
    keys = ("image_size", "model_channels", "out_channels", "num_res_blocks", "attention_resolutions", "channel_mult",
            "num_heads", "num_head_channels", "use_scale_shift_norm", "resblock_updown", "use_new_attention_order")
    net = UA.UNetModel(in_channels=3, **{k: cfg[k] for k in keys}).eval()
    snet = UA.SigmaModel(dim=sg["dim"], channels=sg["channels"], n_blocks=sg["n_blocks"], num_heads=cfg["num_heads"],
                         num_head_channels=cfg["num_head_channels"],
                         use_new_attention_order=cfg["use_new_attention_order"]).eval()
    net.load_state_dict(sd)
    snet.load_state_dict(ssd)
    return cfg, sg, sd, ssd, net, snet


def adm():
    """ADM UNet / sigma-model outputs of the unmodified reference (src/unet_adm.py) -> nets_adm.pt"""
    torch.set_num_threads(4)
    gold = {}
    for name in ("adm_tiny", "adm_alt"):
        cfg, sg, sd, ssd, net, snet = adm_reference_modules(name)
        g = torch.Generator().manual_seed(12)
        x = torch.randn(2, 3, cfg["image_size"], cfg["image_size"], generator=g)
        t = torch.tensor([731.0, 44.0])
        with torch.no_grad():
            out, feat = net(x, t), net.encode(x, t)
            r = snet(feat)
        gold[name] = dict(x=x, t=t, out=out, feat=feat, r=r)
    torch.save(gold, os.path.join(HERE, "nets_adm.pt"))


LOOP2_CASES = [  # (loop, continuous_t, clip, rates, sampler, eta)
    ("denoise", True, "clamp", None, "ddim", 0.0),
    ("denoise", True, "dynamic", None, "ddim_simple_orig", 0.85),
    ("project", True, "dynamic", [0.5, 0.2, 0.2, 0.1], "ddim_simple_orig", 0.85),
    ("project", False, "clamp", [1, 0, 0, 0], "ddim", 0.0),
    ("project", True, "clamp", [0, 1, 0, 0], "ddim", 0.0),
]


def image_sample_module():
    """The reference's driver module (holds the module-level projection_loop); plotting / SSIM imports are stubbed."""
    import importlib
    import types
    refimport.load()
    for name in ("skimage", "skimage.metrics", "cv2"):
        try:
            importlib.import_module(name)
        except Exception:
            sys.modules[name] = types.ModuleType(name)
    if not hasattr(sys.modules["skimage.metrics"], "structural_similarity"):
        sys.modules["skimage.metrics"].structural_similarity = None
        sys.modules["skimage"].metrics = sys.modules["skimage.metrics"]
    return importlib.import_module("image_sample")


def loops2():
    """continuous_t, dynamic thresholding and the sigma feed-forward projection_loop, observed on the reference
    -> loops2_tiny.pt"""
    R = refimport.load()
    IS = image_sample_module()
    torch.set_num_threads(4)
    cfg = weights.CONFIGS["tiny"]
    u, sg = cfg["unet"], cfg["sigma"]
    net = R.unet_ddim.UNetModel(**u).eval()
    net.load_state_dict(weights.ddim_unet_state_dict(**u, seed=3))
    snet = R.unet_ddim.SigmaModel(dim=sg["dim"], channels=sg["channels"], n_blocks=sg["n_blocks"]).eval()
    snet.load_state_dict(weights.ddim_sigma_state_dict(**sg, seed=4))
    B, side = 2, u["image_size"]
    shape = (B, 3, side, side)
    gold = {}
    for loop, cont, clip, rates, kind, eta in LOOP2_CASES:
        sch = R.schedulers.get_sampler(kind, 1000, 6, start_sigma=20.0, eta=eta, continuous_t=cont)
        exp = R.experiments.ImageExperiment(net, sch, batch_size=B, data_shape=shape[1:], seed=5, device="cpu")
        exp.set_model(net, snet, learn_epsvar=False)
        exp.set_norm_maxmin(0.0, 30.0)
        exp.set_clip_fn(clip)
        torch.manual_seed(5)
        z = torch.randn(shape)
        noises = [torch.randn(shape) for _ in range(len(sch.timesteps) - 1)] if eta > 0 else []
        rec = dict(xt=[], sigma_t=[], sigma_prev=[], x_prev=[], x0=[])
        orig = sch.pred_xprev

        def spy(*a, _orig=orig, _rec=rec, **k):
            out = _orig(*a, **k)
            _rec["xt"].append(k["xt"].clone())
            _rec["x0"].append(k["x0"].clone())
            _rec["sigma_t"].append(torch.as_tensor(k["sigma_t"]).reshape(-1).clone())
            _rec["sigma_prev"].append(torch.as_tensor(k["sigma_prev"]).reshape(-1).clone())
            _rec["x_prev"].append(out.clone())
            return out

        sch.pred_xprev = spy
        kw = dict(shape=shape, gen=exp.new_gen(5), style="pred", norm_eps=True, refine_prior_sigma=True, chunk_size=1)
        if loop == "denoise":
            final, _ = exp.denoise_loop(return_log=False, **kw)
        else:
            final, _ = IS.projection_loop(exp, sigma_estimate_rate=rates, **kw)
        gold["%s|%d|%s|%s|%s|%s" % (loop, int(cont), clip, rates, kind, eta)] = dict(
            z=z, noises=noises, final=final, timesteps=sch.timesteps.clone(), sigmas=sch.sampling_sigmas.clone(), **rec)
    torch.save(gold, os.path.join(HERE, "loops2_tiny.pt"))


def operators2():
    """Operators added after the first golden set: the anisotropic blur Deblurring2D with the reference's deblur_aniso
    kernels (src/constraint_functions.py:280-292)  -> operators2_r32.pt"""
    R = refimport.load()
    ref = R.svd_operators
    Rr, C, Bo = 32, 3, 2
    g = torch.Generator().manual_seed(22)
    xs = torch.rand(Bo, C * Rr * Rr, generator=g) * 2 - 1
    x0 = torch.randn(Bo, C, Rr, Rr, generator=g)
    k1, k2 = O.aniso_kernels()
    op = ref.Deblurring2D(k1, k2, C, Rr, "cpu")
    y = op.A(xs.clone())
    proj = x0 - op.A_pinv(op.A(x0.reshape(Bo, -1)) - y).reshape(x0.shape)
    gold = dict(x=xs, x0=x0, deblur_aniso=dict(A=y, At=op.At(y.clone()), A_pinv=op.A_pinv(y.clone()), project=proj))
    torch.save(gold, os.path.join(HERE, "operators2_r32.pt"))


CONSTRAINED_TASKS = [("sr_averagepooling", 4), ("inpainting_box", 1), ("colorization", 1), ("cs_walshhadamard", 4)]


def loops3():
    """DDNM-constrained restoration with NLC (configs c4/c5) on the unmodified reference: ADM UNet with the learned
    variance head (adm_tiny), ddim_simple_orig eta 0.85, dynamic clip, the `svd` projection x0 - A^+(A x0 - y) built
    the way image_sample.py:359-383,636-648 builds it (Constraint_Function over the reference's own svd_operators;
    `get_constraint_function` itself is not callable)  -> loops3_constrained.pt"""
    from functools import partial
    R = refimport.load()
    IS = image_sample_module()
    torch.set_num_threads(4)
    cfg, sg, sd, ssd, net, snet = adm_reference_modules("adm_tiny")
    side, B, C = cfg["image_size"], 2, 3
    shape = (B, C, side, side)
    ref = R.svd_operators
    mask = torch.ones(side, side)
    q = side // 4
    mask[q:3 * q, q:3 * q] = 0
    mr = torch.nonzero(mask.reshape(-1) == 0).long().reshape(-1) * 3
    missing = torch.cat([mr, mr + 1, mr + 2])
    perm = torch.randperm(side * side, generator=torch.Generator().manual_seed(3))
    gold = dict(missing=missing, perm=perm, mask=mask)
    for task, scale in CONSTRAINED_TASKS:
        A_funcs = {"sr_averagepooling": lambda: ref.SuperResolution(C, side, scale, "cpu"),
                   "inpainting_box": lambda: ref.Inpainting(C, side, missing, "cpu"),
                   "colorization": lambda: ref.Colorization(side, "cpu"),
                   "cs_walshhadamard": lambda: ref.WalshHadamardCS(C, side, scale, perm, "cpu")}[task]()
        A, Ap = A_funcs.A, A_funcs.A_pinv

        def affine_svd(x0_t, y, lambda_t, A, Ap):  # image_sample.py:376-379
            return x0_t - Ap(A(x0_t.reshape(x0_t.size(0), -1)) - y.reshape(y.size(0), -1)).reshape(*x0_t.size())

        deg = "inpainting" if task.startswith("inpainting") else task
        con = IS.Constraint_Function(deg, A, Ap, partial(affine_svd, A=A, Ap=Ap), proj="svd", channels=C,
                                     image_size=side, lr=1.0)
        g = torch.Generator().manual_seed(31)
        x_true = torch.rand(shape, generator=g) * 2 - 1
        y = con.transform(x_true)  # image_sample.py:638
        constrain_fn = partial(con.constraint_fn, y=y, lambda_t=con.lr)  # :647
        constrain_loss = partial(con.loss, y=y)  # :648
        sch = R.schedulers.get_sampler("ddim_simple_orig", 1000, 5, start_sigma=20.0, eta=0.85, sampler_var="learned")
        exp = R.experiments.ImageExperiment(net, sch, batch_size=B, data_shape=shape[1:], seed=5, device="cpu")
        exp.set_model(net, snet, learn_epsvar=True)
        exp.set_norm_maxmin(0.0, 30.0)
        exp.set_clip_fn("dynamic")
        torch.manual_seed(5)
        z = torch.randn(shape)
        noises = [torch.randn(shape) for _ in range(len(sch.timesteps) - 1)]
        rec = dict(xt=[], sigma_t=[], sigma_prev=[], x_prev=[], x0=[], eps=[])
        orig = sch.pred_xprev

        def spy(*a, _orig=orig, _rec=rec, **k):
            out = _orig(*a, **k)
            for name in ("xt", "x0", "eps"):
                _rec[name].append(k[name].clone())
            _rec["sigma_t"].append(torch.as_tensor(k["sigma_t"]).reshape(-1).clone())
            _rec["sigma_prev"].append(torch.as_tensor(k["sigma_prev"]).reshape(-1).clone())
            _rec["x_prev"].append(out.clone())
            return out

        sch.pred_xprev = spy
        final, logs = exp.denoise_loop(shape=shape, gen=exp.new_gen(5), style="pred", constrain_fn=constrain_fn,
                                       norm_eps=True, refine_prior_sigma=True, return_log=True, chunk_size=1,
                                       constrain_loss=constrain_loss, sigma_pred_threshold=960)
        gold["%s|%d" % (task, scale)] = dict(
            x_true=x_true, y=y, z=z, noises=noises, final=final, x0_hat=logs[2], const=logs[4],
            timesteps=sch.timesteps.clone(), sigmas=sch.sampling_sigmas.clone(), **rec)
    torch.save(gold, os.path.join(HERE, "loops3_constrained.pt"))


EDM_CASES = [("pred_partial,pred", "00", False, 1.0), ("base,base", "00", False, 1.0), ("pred,pred_partial", "11", True, 1.0),
             ("pred_sigma,pred_partial3", "10", False, None), ("pred_partial,pred", "01", True, 1.004)]


def edm_reference_modules(name):
    import importlib
    refimport.load()
    EN = importlib.import_module("src.edm_networks")
    cfg = dict(weights.EDM_CONFIGS[name])
    sg = cfg.pop("sigma")
    sd = weights.edm_unet_state_dict(**cfg, seed=3)
    ssd = weights.edm_sigma_state_dict(**sg, seed=4)
    net = EN.SongUNet(**{k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.items()}).eval()
    snet = EN.SigmaModel(dim=sg["dim"], channels=sg["channels"], n_blocks=sg["n_blocks"]).eval()
    net.load_state_dict(sd)
    snet.load_state_dict(ssd)
    return cfg, sg, sd, ssd, net, snet


def edm():
    """EDM network outputs and the reference's own edm_sampler (4 steps = 7 NFE), with every get_denoise_vector
    call observed -> nets_edm.pt, edm_sampler_tiny.pt"""
    R = refimport.load()
    torch.set_num_threads(4)
    cfg, sg, sd, ssd, net, snet = edm_reference_modules("edm_tiny")
    Rr = cfg["img_resolution"]
    g = torch.Generator().manual_seed(13)
    x = torch.randn(2, 3, Rr, Rr, generator=g)
    c_noise = torch.tensor([0.91, -1.35])
    with torch.no_grad():
        out, feat = net(x, c_noise, None), net.encode(x, c_noise, None)
        r = snet(feat)
    torch.save(dict(x=x, c_noise=c_noise, out=out, feat=feat, r=r), os.path.join(HERE, "nets_edm.pt"))

    B = 2
    gold = {}
    for style, ne, refine, es in EDM_CASES:
        exp = R.experiments.EDMImageExperiment(net, None, batch_size=B, data_shape=(3, Rr, Rr), seed=1, device="cpu",
                                               num_timesteps=4, sigma_min=0.002, sigma_max=80)
        exp.set_model(net, snet, learn_epsvar=False)
        exp.set_norm_maxmin(0.0, 30.0)
        gen = R.experiments.StackedRandomGenerator("cpu", [0, 1])
        latents = gen.randn((B, 3, Rr, Rr), device="cpu")
        gen = R.experiments.StackedRandomGenerator("cpu", [0, 1])
        calls = []
        orig = exp.get_denoise_vector

        def spy(xt, sigma_t, sigma_prev, _orig=orig, _calls=calls, **k):
            res = _orig(xt, sigma_t, sigma_prev, **k)
            f = lambda v: torch.as_tensor(v).reshape(-1).double().clone()
            _calls.append(dict(xt=xt.clone(), sigma_in=f(sigma_t), sigma_prev_in=f(sigma_prev), style=k["style"],
                               eps=res[0].clone(), sigma_t=f(res[2]), sigma_prev=f(res[3])))
            return res

        exp.get_denoise_vector = spy
        with torch.no_grad():
            final = exp.edm_sampler((B, 3, Rr, Rr), gen=gen, style=style, norm_eps=ne + "0", refine_prior_sigma=refine,
                                    eps_scale=es)
        gold["%s|%s|%d|%s" % (style, ne, int(refine), es)] = dict(latents=latents, final=final, calls=calls)
    torch.save(gold, os.path.join(HERE, "edm_sampler_tiny.pt"))


DDNM_REGIMES = [(0.9, 0.1, 0.05, 0.85), (0.9, 0.1, 0.3, 0.85), (0.9, 0.1, 0.6, 0.5), (0.9, 0.0, 0.3, 0.85),
                (0.3, 0.5, 0.12, 0.0)]  # (a, sigma_y, sigma_t, eta): both sides of sigma_t <> a sigma_y / s, and sigma_y = 0
DDNM_LOOPS = [("colorization", 0.1), ("sr_averagepooling", 0.2), ("inpainting", 0.05), ("cs_walshhadamard", 0.1),
              ("deblur_gauss", 0.05), ("denoising", 0.3), ("sr_averagepooling", None), ("deblur_gauss", None)]


def _ddnm_ops(ref, Rr, C, missing, perm):
    return {"inpainting": ref.Inpainting(C, Rr, missing, "cpu"), "colorization": ref.Colorization(Rr, "cpu"),
            "sr_averagepooling": ref.SuperResolution(C, Rr, 4, "cpu"),
            "cs_walshhadamard": ref.WalshHadamardCS(C, Rr, 4, perm, "cpu"),
            "deblur_gauss": ref.Deblurring(O.gauss_kernel(), C, Rr, "cpu"), "denoising": ref.Denoising(C, Rr, "cpu")}


def ddnm():
    """DDNM+ on the unmodified reference: Lambda / Lambda_noise / A_pinv_eta of every operator
    class that defines them (functions/svd_operators.py) in five (a, sigma_y, sigma_t, eta) regimes -> ddnm_ops_r32.pt, and
    per-step dumps of functions/svd_ddnm.py ddnm_diffusion / ddnm_plus_diffusion (adm_tiny network, 4 sampling steps with
    one time-travel detour) -> ddnm_loops_r32.pt.  The loops hard-code `.to('cuda')`; the only change made here is to redirect
    those moves to the CPU while the loop runs."""
    import importlib
    import types
    R = refimport.load()
    ref = R.svd_operators
    sys.modules.setdefault("torchvision", types.ModuleType("torchvision"))
    if not hasattr(sys.modules["torchvision"], "utils"):
        sys.modules["torchvision"].utils = types.ModuleType("torchvision.utils")
        sys.modules["torchvision.utils"] = sys.modules["torchvision"].utils
    SD = importlib.import_module("functions.svd_ddnm")
    torch.set_num_threads(4)
    Rr, C, Bo = 32, 3, 2
    g = torch.Generator().manual_seed(41)
    v = torch.randn(Bo, C * Rr * Rr, generator=g)
    e = torch.randn(Bo, C * Rr * Rr, generator=g)
    mask = torch.ones(Rr, Rr)
    mask[8:24, 8:24] = 0
    mr = torch.nonzero(mask.reshape(-1) == 0).long().reshape(-1) * 3
    missing = torch.cat([mr, mr + 1, mr + 2])
    perm = torch.randperm(Rr * Rr, generator=torch.Generator().manual_seed(3))
    ops = _ddnm_ops(ref, Rr, C, missing, perm)
    gold = dict(v=v, e=e, missing=missing, perm=perm, regimes=DDNM_REGIMES)
    for name, op in ops.items():
        rec = dict(Lambda=[], Lambda_noise=[])
        for a, sy, st, eta in DDNM_REGIMES:
            a_t, st_t = torch.tensor(a), torch.tensor(st)  # 0-dim fp32, as functions/svd_ddnm.py:121-132 passes them
            rec["Lambda"].append(op.Lambda(v.clone(), a_t, sy, st_t, eta))
            rec["Lambda_noise"].append(op.Lambda_noise(v.clone(), a_t, sy, st_t, eta, e.clone()))
        if name != "denoising":
            yy = op.A(v.clone())
            rec["y"] = yy
            rec["A_pinv_eta"] = [op.A_pinv_eta(yy.clone(), 0.01), op.A_pinv_eta(yy.clone(), 0.5)]
        gold[name] = rec
    torch.save(gold, os.path.join(HERE, "ddnm_ops_r32.pt"))

    # ---- loops
    cfg, sg, sd, ssd, net, snet = adm_reference_modules("adm_tiny")
    betas = torch.linspace(1e-4, 2e-2, 1000)  # the DDNM configs' linear schedule
    config = types.SimpleNamespace(diffusion=types.SimpleNamespace(num_diffusion_timesteps=1000),
                                   time_travel=types.SimpleNamespace(T_sampling=4, travel_length=2, travel_repeat=2))
    loops = dict(betas=betas, T_sampling=4, travel_length=2, travel_repeat=2, eta=0.85, missing=missing, perm=perm)
    orig_to, orig_randn_like = torch.Tensor.to, torch.randn_like

    def to_cpu(self, *a, **k):
        a = tuple("cpu" if (isinstance(x, str) and x.startswith("cuda")) else x for x in a)
        return orig_to(self, *a, **k)

    for name, sigma_y in DDNM_LOOPS:
        op = ops[name]
        gg = torch.Generator().manual_seed(51)
        x_true = torch.rand(Bo, C, Rr, Rr, generator=gg) * 2 - 1
        y = op.A(x_true.reshape(Bo, -1).clone())
        if sigma_y:
            y = y + sigma_y * torch.randn(y.shape, generator=gg)
        xT = torch.randn(Bo, C, Rr, Rr, generator=gg)
        rec = dict(xt=[], t=[], et=[], z=[])

        def model(x, t, _rec=rec):
            out = net(x, t)
            _rec["xt"].append(x.clone()), _rec["t"].append(t.clone()), _rec["et"].append(out[:, :3].clone())
            return out

        def randn_like(x, _rec=rec, _g=gg):
            z = torch.randn(x.shape, generator=_g)
            _rec["z"].append(z.clone())
            return z

        torch.Tensor.to, torch.randn_like = to_cpu, randn_like
        try:
            if sigma_y is None:
                xs, x0s = SD.ddnm_diffusion(xT.clone(), model, betas, 0.85, op, y, config=config)
            else:
                xs, x0s = SD.ddnm_plus_diffusion(xT.clone(), model, betas, 0.85, op, y, sigma_y, config=config)
        finally:
            torch.Tensor.to, torch.randn_like = orig_to, orig_randn_like
        loops["%s|%s" % (name, sigma_y)] = dict(x_true=x_true, y=y, xT=xT, x_last=xs[0], x0_last=x0s[0], **rec)
    torch.save(loops, os.path.join(HERE, "ddnm_loops_r32.pt"))
    for f in ("ddnm_ops_r32.pt", "ddnm_loops_r32.pt"):
        print(f, os.path.getsize(os.path.join(HERE, f)))


def operators3():
    """Block-wise CS (functions/svd_operators.py:101-160, with its random basis replaced by oracle.hadamard_basis so that
    the fixture does not have to store a 1024 x 1024 matrix) and GeneralA (:173-208) on a stored 40 x 96 matrix
    -> operators3.pt"""
    R = refimport.load()
    ref = R.svd_operators
    g = torch.Generator().manual_seed(23)
    Rr, C, Bo = 64, 3, 2
    op = ref.CS(C, Rr, 0.25, "cpu")
    op.V_small = O.hadamard_basis(1024, 7)
    op.Vt_small = op.V_small.transpose(0, 1)
    xs = torch.rand(Bo, C * Rr * Rr, generator=g) * 2 - 1
    x0 = torch.randn(Bo, C, Rr, Rr, generator=g)
    y = op.A(xs.clone())
    proj = x0 - op.A_pinv(op.A(x0.reshape(Bo, -1)) - y).reshape(x0.shape)
    gold = dict(cs=dict(x=xs, x0=x0, A=y, At=op.At(y.clone()), A_pinv=op.A_pinv(y.clone()),
                        A_pinv_eta=op.A_pinv_eta(y.clone(), 0.1), project=proj))
    Am = torch.randn(40, 96, generator=g)
    ga = ref.GeneralA(Am.clone())
    xg = torch.randn(3, 96, generator=g)
    x0g = torch.randn(3, 96, generator=g)
    yg = ga.A(xg.clone())
    gold["general"] = dict(Amat=Am, x=xg, x0=x0g, A=yg, At=ga.At(yg.clone()), A_pinv=ga.A_pinv(yg.clone()),
                           A_pinv_eta=ga.A_pinv_eta(yg.clone(), 0.1),
                           project=x0g - ga.A_pinv(ga.A(x0g.clone()) - yg))
    torch.save(gold, os.path.join(HERE, "operators3.pt"))
    print("operators3.pt", os.path.getsize(os.path.join(HERE, "operators3.pt")))


def basicsr_psnr_ssim():
    """basicsr.metrics.psnr_ssim of the reference with its unused skimage import stubbed."""
    import importlib
    import types
    refimport.load()
    for name in ("skimage", "skimage.metrics"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if not hasattr(sys.modules["skimage.metrics"], "structural_similarity"):
        sys.modules["skimage.metrics"].structural_similarity = None
        sys.modules["skimage"].metrics = sys.modules["skimage.metrics"]
    return importlib.import_module("basicsr.metrics.psnr_ssim")


def reference_ssim_fn(PS, sample, orig):
    """image_sample.py:571-582 on the CPU: basicsr's `_ssim_3d` hard-codes `.cuda()`; those moves are redirected."""
    orig_t, orig_m = torch.Tensor.cuda, torch.nn.Module.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self
    try:
        sample = torch.round(sample * 255).to(torch.uint8)
        orig = torch.round(orig * 255).to(torch.uint8)
        return [float(PS.calculate_ssim(sample[i], orig[i], crop_border=0, test_y_channel=False))
                for i in range(len(sample))]
    finally:
        torch.Tensor.cuda, torch.nn.Module.cuda = orig_t, orig_m


def ssim():
    """SSIM as the reference's evaluation computes it (image_sample.py:571-582 -> basicsr calculate_ssim, ssim3d) for
    noisy / smooth / identical image pairs at 32 x 32 and a ragged 40 x 52 -> ssim.pt"""
    PS = basicsr_psnr_ssim()
    g = torch.Generator().manual_seed(61)
    gold = {}
    for name, (H, W) in (("r32", (32, 32)), ("ragged", (40, 52))):
        smooth = torch.nn.functional.avg_pool2d(torch.rand(4, 3, H + 4, W + 4, generator=g), 5, 1)
        noisy = (smooth + 0.05 * torch.randn(4, 3, H, W, generator=g)).clamp(0, 1)
        rnd = torch.rand(4, 3, H, W, generator=g)
        cases = dict(noisy_vs_smooth=(noisy, smooth), random_vs_smooth=(rnd, smooth), same=(noisy, noisy.clone()))
        gold[name] = {k: dict(sample=a, orig=b, ssim=torch.tensor(reference_ssim_fn(PS, a, b), dtype=torch.float64))
                      for k, (a, b) in cases.items()}
    torch.save(gold, os.path.join(HERE, "ssim.pt"))
    print("ssim.pt", os.path.getsize(os.path.join(HERE, "ssim.pt")))


def train_step():
    """One sigma-model training iteration of the reference (src/experiments.py:683-694) on its own modules: the DDIM
    SigmaModel (src/unet_ddim.py:493-529, dropout 0) in train() mode on recorded encoder features, MSE against a target
    noise level, backward, torch.optim.AdamW step and the EMA update -> train_step_tiny.pt (loss, every parameter's gradient,
    the updated parameters and EMA copy).  The targets for a native backward pass."""
    import copy
    R = refimport.load()
    torch.set_num_threads(4)
    cfg = weights.CONFIGS["tiny"]
    ssd = weights.ddim_sigma_state_dict(**cfg["sigma"], seed=4)
    net = R.unet_ddim.SigmaModel(dropout=0.0, **cfg["sigma"])
    net.load_state_dict(ssd)
    net.train()
    g = torch.Generator().manual_seed(71)
    feat = torch.load(os.path.join(HERE, "nets_tiny.pt"), weights_only=True)["feat"]
    feat = torch.cat([feat, feat.flip(0) * 0.7 + 0.1 * torch.randn(feat.shape, generator=g)])  # batch 4 for BatchNorm
    dist_real = (1.0 + 0.2 * torch.randn(feat.shape[0], 1, 1, 1, generator=g))
    optim = torch.optim.AdamW(net.parameters(), lr=1e-3, weight_decay=0.01)
    ema = copy.deepcopy(list(net.parameters()))
    dist_hat = net(feat) + 1
    loss = torch.nn.MSELoss()(dist_real, dist_hat)
    optim.zero_grad()
    loss.backward()
    grads = {n: p.grad.clone() for n, p in net.named_parameters()}
    optim.step()
    R.experiments  # (update_ema is src/nn_util.py:55-65: targ.mul_(rate).add_(src, alpha=1-rate))
    for targ, src in zip(ema, net.parameters()):
        targ.detach().mul_(0.999).add_(src.detach(), alpha=1 - 0.999)
    # the sigma-model has 3.9 M parameters: every tensor is stored as (L2 norm, sum, first 32 entries), the small head
    # layers in full
    def digest(t):
        t = t.detach().double().reshape(-1)
        return dict(norm=t.norm().clone(), sum=t.sum().clone(), head=t[:32].clone(), full=t.float().clone() if t.numel() <= 4096 else None)

    gold = dict(feat=feat, dist_real=dist_real, dist_hat=dist_hat.detach(), loss=loss.detach(),
                grads={n: digest(v) for n, v in grads.items()},
                new_params={n: digest(p) for n, p in net.named_parameters()},
                ema={n: digest(e) for (n, _), e in zip(net.named_parameters(), ema)})
    torch.save(gold, os.path.join(HERE, "train_step_tiny.pt"))
    print("train_step_tiny.pt", os.path.getsize(os.path.join(HERE, "train_step_tiny.pt")))


# ---------------------------------------------------------------------------------------------------------------------
# Benchmark architectures (BASELINE.json configs c2, c3, c4/c5): the unmodified reference at its REAL sizes.
C2_SNAPSHOTS = (0, 3, 10, 25, 50, 75, 90, 99)  # steps whose tensors are stored for teacher-forced checks


def _checksum(t):
    """Position-weighted float64 digest: pins a tensor that the test regenerates from its seed instead of loading."""
    v = t.double().reshape(-1)
    return torch.stack([v.sum(), (v * torch.arange(1, v.numel() + 1, dtype=torch.float64)).sum() / v.numel()])


def loop_c2():
    """Config c2 on the unmodified reference: CelebA-64 unet_ddim + sigma-model, ddim_simple_orig eta 0.85, 100 steps
    from sigma 100, style pred + norm_eps + refine, clamp clip, norm_min -2 / norm_max 110 (bench.py's c2), batch 4.
    Stored: the final image, every step's corrected sigmas and times, full tensors at C2_SNAPSHOTS.  The initial and
    per-step noise is NOT stored (100 x 196 KB): the test redraws it from seed 5 and checks the digests  -> loop_c2_100.pt"""
    R = refimport.load()
    torch.set_num_threads(8)
    cfg = weights.CONFIGS["c2"]
    u, sg = cfg["unet"], cfg["sigma"]
    net = R.unet_ddim.UNetModel(**u).eval()
    net.load_state_dict(weights.ddim_unet_state_dict(**u, seed=3))
    snet = R.unet_ddim.SigmaModel(dim=sg["dim"], channels=sg["channels"], n_blocks=sg["n_blocks"]).eval()
    snet.load_state_dict(weights.ddim_sigma_state_dict(**sg, seed=4))
    B, side, steps = 4, u["image_size"], 100
    shape = (B, 3, side, side)
    sch = R.schedulers.get_sampler("ddim_simple_orig", 1000, steps, start_sigma=100, eta=0.85)
    exp = R.experiments.ImageExperiment(net, sch, batch_size=B, data_shape=shape[1:], seed=5, device="cpu")
    exp.set_model(net, snet, learn_epsvar=False)
    exp.set_norm_maxmin(-2.0, 110.0)
    exp.set_clip_fn("clamp")
    torch.manual_seed(5)
    z = torch.randn(shape)
    noises = [torch.randn(shape) for _ in range(len(sch.timesteps) - 1)]
    rec = dict(sigma_t=[], sigma_prev=[], snap={})
    orig = sch.pred_xprev

    def spy(*a, _orig=orig, **k):
        out = _orig(*a, **k)
        i = len(rec["sigma_t"])
        rec["sigma_t"].append(torch.as_tensor(k["sigma_t"]).reshape(-1).clone())
        rec["sigma_prev"].append(torch.as_tensor(k["sigma_prev"]).reshape(-1).clone())
        if i in C2_SNAPSHOTS:
            rec["snap"][i] = dict(xt=k["xt"].clone(), eps=k["eps"].clone(), x0=k["x0"].clone(), x_prev=out.clone())
        return out

    sch.pred_xprev = spy
    # the two time lookups of every step (observed, not modified): t of the refined sigma for the encode pass, t_hat of
    # the corrected sigma for the forward pass (src/experiments.py:410,427)
    t_first, t_second = [], []
    orig_enc, orig_pred = exp.encode_xt, exp.pred_xt

    def enc_spy(xt, t, sigma=None, batch_t=True):
        t_first.append(torch.as_tensor(t).reshape(-1).float().expand(B).clone())
        return orig_enc(xt, t, sigma=sigma, batch_t=batch_t)

    def pred_spy(xt, t, sigma=None, batch_t=True):
        t_second.append(torch.as_tensor(t).reshape(-1).float().expand(B).clone())
        return orig_pred(xt, t, sigma=sigma, batch_t=batch_t)

    exp.encode_xt, exp.pred_xt = enc_spy, pred_spy
    final, _ = exp.denoise_loop(shape=shape, gen=exp.new_gen(5), style="pred", norm_eps=True, refine_prior_sigma=True,
                                return_log=False, chunk_size=1, sigma_pred_threshold=960)
    sig_hat = torch.stack([v.expand(B) for v in rec["sigma_t"]])  # [steps, B] (scalar on the 'base' steps)
    t_hat = torch.searchsorted(sch.sigmas, sig_hat.contiguous())  # the time bucket each corrected sigma falls in
    assert len(t_first) == steps and len(t_second) == steps and torch.equal(torch.stack(t_second).long(), t_hat)
    torch.save(dict(final=final, t_first=torch.stack(t_first), z_digest=_checksum(z), noise_digest=torch.stack([_checksum(n) for n in noises]),
                    sigma_t=sig_hat, sigma_prev=torch.stack([v.expand(B) for v in rec["sigma_prev"]]), t_hat=t_hat,
                    snap=rec["snap"], timesteps=sch.timesteps.clone(), sigmas=sch.sampling_sigmas.clone(),
                    table=sch.sigmas.clone()), os.path.join(HERE, "loop_c2_100.pt"))


def nets_bench():
    """Network outputs of the unmodified reference at the c3 (EDM SongUNet-64) and c4/c5 (ADM-256) architectures
    -> nets_bench.pt"""
    torch.set_num_threads(8)
    gold = {}
    cfg, sg, sd, ssd, net, snet = edm_reference_modules("edm64")
    Rr = cfg["img_resolution"]
    g = torch.Generator().manual_seed(14)
    x = torch.randn(2, 3, Rr, Rr, generator=g)
    c_noise = torch.tensor([0.73, -1.1])
    with torch.no_grad():
        out, feat = net(x, c_noise, None), net.encode(x, c_noise, None)
        r = snet(feat)
    gold["edm64"] = dict(x=x, c_noise=c_noise, out=out, feat=feat, r=r)
    del net, snet, sd, ssd
    cfg, sg, sd, ssd, net, snet = adm_reference_modules("adm256")
    del sd, ssd
    g = torch.Generator().manual_seed(15)
    x = torch.randn(1, 3, cfg["image_size"], cfg["image_size"], generator=g)
    t = torch.tensor([412.0])
    with torch.no_grad():
        out, feat = net(x, t), net.encode(x, t)
        r = snet(feat)
    gold["adm256"] = dict(x=x, t=t, out=out, feat=feat, r=r)
    torch.save(gold, os.path.join(HERE, "nets_bench.pt"))


def steps_adm256():
    """Configs c4 / c5 on the unmodified reference at 256 x 256: the DDNM-constrained NLC loop (ADM-256 with the learned
    variance head, ddim_simple_orig eta 0.85, dynamic clip, svd projection) for SR x4 (c4) and colourisation (c5),
    batch 1, a 2-step schedule from sigma 20 (3 loop iterations incl. the final one to sigma 0), every pred_xprev call
    observed  -> steps_adm256.pt"""
    from functools import partial
    R = refimport.load()
    IS = image_sample_module()
    torch.set_num_threads(8)
    cfg, sg, sd, ssd, net, snet = adm_reference_modules("adm256")
    del sd, ssd
    side, B, C = cfg["image_size"], 1, 3
    shape = (B, C, side, side)
    ref = R.svd_operators
    gold = {}
    for task, scale in (("sr_averagepooling", 4), ("colorization", 1)):
        A_funcs = ref.SuperResolution(C, side, scale, "cpu") if task == "sr_averagepooling" else ref.Colorization(side, "cpu")
        A, Ap = A_funcs.A, A_funcs.A_pinv

        def affine_svd(x0_t, y, lambda_t, A, Ap):  # image_sample.py:376-379
            return x0_t - Ap(A(x0_t.reshape(x0_t.size(0), -1)) - y.reshape(y.size(0), -1)).reshape(*x0_t.size())

        con = IS.Constraint_Function(task, A, Ap, partial(affine_svd, A=A, Ap=Ap), proj="svd", channels=C,
                                     image_size=side, lr=1.0)
        g = torch.Generator().manual_seed(32)
        x_true = torch.rand(shape, generator=g) * 2 - 1
        y = con.transform(x_true)
        sch = R.schedulers.get_sampler("ddim_simple_orig", 1000, 2, start_sigma=20.0, eta=0.85, sampler_var="learned")
        exp = R.experiments.ImageExperiment(net, sch, batch_size=B, data_shape=shape[1:], seed=5, device="cpu")
        exp.set_model(net, snet, learn_epsvar=True)
        exp.set_norm_maxmin(-2.0, 110.0)
        exp.set_clip_fn("dynamic")
        torch.manual_seed(5)
        z = torch.randn(shape)
        noises = [torch.randn(shape) for _ in range(len(sch.timesteps) - 1)]
        rec = dict(sigma_t=[], sigma_prev=[], x_prev=[], x0=[])
        orig = sch.pred_xprev

        def spy(*a, _orig=orig, _rec=rec, **k):
            out = _orig(*a, **k)
            _rec["x0"].append(k["x0"].clone())
            _rec["sigma_t"].append(torch.as_tensor(k["sigma_t"]).reshape(-1).clone())
            _rec["sigma_prev"].append(torch.as_tensor(k["sigma_prev"]).reshape(-1).clone())
            _rec["x_prev"].append(out.clone())
            return out

        sch.pred_xprev = spy
        final, logs = exp.denoise_loop(shape=shape, gen=exp.new_gen(5), style="pred",
                                       constrain_fn=partial(con.constraint_fn, y=y, lambda_t=con.lr), norm_eps=True,
                                       refine_prior_sigma=True, return_log=True, chunk_size=1,
                                       constrain_loss=partial(con.loss, y=y), sigma_pred_threshold=960)
        gold["%s|%d" % (task, scale)] = dict(
            y=y, x_true_digest=_checksum(x_true), z_digest=_checksum(z),
            noise_digest=torch.stack([_checksum(n) for n in noises]), final=final, const=logs[4],
            timesteps=sch.timesteps.clone(), sigmas=sch.sampling_sigmas.clone(), **rec)
    torch.save(gold, os.path.join(HERE, "steps_adm256.pt"))


def dhariwal():
    """DhariwalUNet (src/edm_networks.py:406-502) outputs of the unmodified reference: a two-level test shape and the
    64 x 64 four-level shape  -> nets_dhariwal.pt"""
    import importlib
    refimport.load()
    EN = importlib.import_module("src.edm_networks")
    torch.set_num_threads(8)
    gold = {}
    for name in ("dhariwal_tiny", "dhariwal64"):
        cfg = dict(weights.DHARIWAL_CONFIGS[name])
        net = EN.DhariwalUNet(**{k: (list(v) if isinstance(v, tuple) else v) for k, v in cfg.items()}).eval()
        net.load_state_dict(weights.dhariwal_unet_state_dict(**cfg, seed=3))
        g = torch.Generator().manual_seed(16)
        x = torch.randn(2, 3, cfg["img_resolution"], cfg["img_resolution"], generator=g)
        c_noise = torch.tensor([0.4, -0.9])
        with torch.no_grad():
            gold[name] = dict(x=x, c_noise=c_noise, out=net(x, c_noise, None))
    torch.save(gold, os.path.join(HERE, "nets_dhariwal.pt"))


def simple_feat_layer():
    """src/unet_simple.py Model(config) with config.model.feat_layer = 1 (the DDIM UNet the factory builds): encode /
    forward_and_encode of the unmodified reference on the tiny architecture -> nets_simple_fl1.pt"""
    import importlib
    import types
    refimport.load()
    US = importlib.import_module("src.unet_simple")
    u = weights.CONFIGS["tiny"]["unet"]
    cfg = types.SimpleNamespace(
        model=types.SimpleNamespace(ch=u["model_channels"], out_ch=u["out_channels"], ch_mult=list(u["channel_mult"]),
                                    num_res_blocks=u["num_res_blocks"], attn_resolutions=list(u["attention_resolutions"]),
                                    dropout=0.0, in_channels=u["in_channels"], resamp_with_conv=True, type="simple",
                                    feat_layer=1),
        data=types.SimpleNamespace(image_size=u["image_size"]),
        diffusion=types.SimpleNamespace(num_diffusion_timesteps=1000))
    net = US.Model(cfg).eval()
    net.load_state_dict(weights.ddim_unet_state_dict(**u, seed=3))
    x = torch.randn(2, 3, u["image_size"], u["image_size"], generator=torch.Generator().manual_seed(8))
    t = torch.tensor([640.0, 12.0])
    with torch.no_grad():
        out, feat = net.forward_and_encode(x, t)
        assert torch.equal(feat, net.encode(x, t))
    torch.save(dict(x=x, t=t, out=out, feat=feat), os.path.join(HERE, "nets_simple_fl1.pt"))


def fid():
    """FID Inception features of a seeded batch, computed with torchvision's own Inception3 modules carrying pytorch_fid's
    pooling patches (oracle.fid.torchvision_fid_inception; pytorch_fid itself is not available) -> fid_tiny.pt"""
    from oracle import fid as OF
    torch.set_num_threads(8)
    sd = weights.fid_inception_state_dict(seed=7)
    g = torch.Generator().manual_seed(21)
    x = torch.rand(2, 3, 48, 48, generator=g)
    samples = torch.randn(2, 3, 64, 64, generator=g) * 0.6  # sampler outputs: some values outside [-1, 1]
    net = OF.torchvision_fid_inception(sd)
    with torch.no_grad():
        feats = net(x)
        feats_s = net(OF.png_round_trip((samples + 1) / 2))
    torch.save(dict(x=x, samples=samples, features=feats, features_of_samples=feats_s), os.path.join(HERE, "fid_tiny.pt"))


def bench_arch():
    loop_c2()
    nets_bench()
    steps_adm256()


if __name__ == "__main__":
    if len(sys.argv) > 1:
        globals()[sys.argv[1]]()
    else:
        main()
