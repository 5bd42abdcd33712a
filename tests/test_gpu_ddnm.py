"""GPU parity of the DDNM+ path through the C ABI (nlc_op_lambda, nlc_op_lambda_noise,
nlc_op_Apinv_eta, nlc_ddnm_step, nlc_ddnm_renoise):

* against the unmodified reference's golden outputs at R = 32 (tests/golden/ddnm_ops_r32.pt: Lambda / Lambda_noise /
  A_pinv_eta of six operator classes in five (a, sigma_y, sigma_t, eta) regimes) — tolerance 1e-5 of max|ref| (fp32; the
  kernels apply V, V^T in closed form, so the summation order differs from the reference's matmul chains);
* against per-step dumps of the reference's own ddnm_diffusion / ddnm_plus_diffusion loops (ddnm_loops_r32.pt):
  teacher-forced on the recorded network outputs (every x_t, the final x_t and x0_t: 2e-5 of max|ref|) and free-running
  with the ADM network in fp32 mode (1e-4, the north-star gate);
* against the oracle on seeded inputs at R = 64, and through size-independent properties at R = 256 (Lambda is the
  identity without measurement noise, Lambda_noise preserves the per-sample norm of d1 v + d2 e, a DDNM step lands
  on the measurement: A x0_hat = y)."""
import math
import os
import types

import pytest
import torch

from oracle import ddnm as OD
from oracle import operators as O
from oracle import weights

pytestmark = pytest.mark.gpu
dev = torch.device("cuda:0")
NAMES = ["inpainting", "colorization", "sr_averagepooling", "cs_walshhadamard", "deblur_gauss", "denoising"]


def _build(name, R, C, missing, perm, ratio=4):
    from nlc_b200 import svd_operators as P
    return {"inpainting": lambda: P.Inpainting(C, R, missing, dev), "colorization": lambda: P.Colorization(R, dev),
            "sr_averagepooling": lambda: P.SuperResolution(C, R, ratio, dev),
            "cs_walshhadamard": lambda: P.WalshHadamardCS(C, R, ratio, perm, dev),
            "deblur_gauss": lambda: P.Deblurring(O.gauss_kernel(), C, R, dev),
            "denoising": lambda: P.Denoising(C, R, dev)}[name]()


def _oracle(name, R, C, missing, perm, ratio=4):
    return {"inpainting": lambda: O.Inpainting(C, R, missing), "colorization": lambda: O.Colorization(R),
            "sr_averagepooling": lambda: O.SuperResolution(C, R, ratio),
            "cs_walshhadamard": lambda: O.WalshHadamardCS(C, R, ratio, perm),
            "deblur_gauss": lambda: O.Deblurring(O.gauss_kernel(), C, R), "denoising": lambda: O.Denoising(C, R)}[name]()


def _close(got, ref, tol, what):
    err = (got.cpu() - ref).abs().max().item()
    assert err <= tol * ref.abs().max().item(), (what, err, ref.abs().max().item())


@pytest.mark.parametrize("name", NAMES)
def test_terms_against_reference_golden(golden_dir, name):
    g = torch.load(os.path.join(golden_dir, "ddnm_ops_r32.pt"), weights_only=True)
    op = _build(name, 32, 3, g["missing"], g["perm"])
    v, e = g["v"].to(dev), g["e"].to(dev)
    for k, (a, sy, st, eta) in enumerate(g["regimes"]):
        _close(op.Lambda(v, a, sy, st, eta), g[name]["Lambda"][k], 1e-5, (name, "Lambda", k))
        _close(op.Lambda_noise(v, a, sy, st, eta, e), g[name]["Lambda_noise"][k], 1e-5, (name, "Lambda_noise", k))
    if name != "denoising":
        for k, eta in enumerate((0.01, 0.5)):
            _close(op.A_pinv_eta(g[name]["y"].to(dev), eta), g[name]["A_pinv_eta"][k], 1e-5, (name, "A_pinv_eta", k))


def test_classes_without_lambda_raise():
    from nlc_b200 import svd_operators as P
    v = torch.randn(1, 3 * 32 * 32, device=dev)
    for op in (P.SRConv(O.bicubic_kernel(4), 3, 32, dev, stride=4), P.Deblurring2D(*O.aniso_kernels(), 3, 32, dev)):
        with pytest.raises(NotImplementedError):
            op.Lambda(v, 0.9, 0.1, 0.3, 0.85)
        with pytest.raises(NotImplementedError):
            op.Lambda_noise(v, 0.9, 0.1, 0.3, 0.85, v)
        # the plain DDNM step only needs A and A^+ and is available
        y = op.A(v)
        x0, xn = op.ddnm_step(v, v * 0.1, v * 0.2, y, 0.5, 0.6, 0.85, None)
        assert torch.isfinite(xn).all()


def _config(g):
    return types.SimpleNamespace(diffusion=types.SimpleNamespace(num_diffusion_timesteps=1000),
                                 time_travel=types.SimpleNamespace(T_sampling=g["T_sampling"],
                                                                   travel_length=g["travel_length"],
                                                                   travel_repeat=g["travel_repeat"]))


def _loop_cases(golden_dir):
    g = torch.load(os.path.join(golden_dir, "ddnm_loops_r32.pt"), weights_only=True)
    return g, [k for k in g if "|" in k]


def test_loops_teacher_forced(golden_dir):
    """The reference's recorded network outputs are replayed, so every difference comes from the fused step kernels."""
    from nlc_b200 import svd_ddnm as SD
    assert SD.get_schedule_jump(4, 2, 2) == [3, 2, 1, 0, 1, 2, 1, 0, -1]
    assert SD.get_schedule_jump(10, 3, 3) == OD.schedule_jump(10, 3, 3)
    g, keys = _loop_cases(golden_dir)
    for key in keys:
        case = g[key]
        name, sy = key.split("|")
        op = _build(name, 32, 3, g["missing"], g["perm"])
        ets, zs, seen = iter(case["et"]), iter(case["z"]), []

        def model(x, t, _ets=ets, _seen=seen):
            _seen.append((x.clone(), t.clone()))
            return next(_ets).to(dev)

        fn = SD.ddnm_diffusion if sy == "None" else SD.ddnm_plus_diffusion
        args = () if sy == "None" else (float(sy),)
        xs, x0s = fn(case["xT"].to(dev), model, g["betas"].to(dev), g["eta"], op, case["y"].to(dev), *args,
                     config=_config(g), noise_fn=lambda like, _z=zs: next(_z).to(dev))
        assert len(seen) == len(case["xt"])
        for k, (x, t) in enumerate(seen):
            assert torch.equal(t.cpu(), case["t"][k]), (key, k)
            _close(x, case["xt"][k], 2e-5, (key, "xt", k))
        _close(xs[0], case["x_last"], 2e-5, (key, "x_last"))
        _close(x0s[0], case["x0_last"], 2e-5, (key, "x0_last"))
        assert xs[0].device.type == "cpu" and x0s[0].device.type == "cpu"  # as the reference returns them


KEYS = ("image_size", "model_channels", "out_channels", "num_res_blocks", "attention_resolutions", "channel_mult",
        "num_heads", "num_head_channels", "use_scale_shift_norm", "resblock_updown", "use_new_attention_order")


@pytest.mark.parametrize("prec,tol", [("fp32", 1e-4), ("bf16", None)])
def test_loops_free_running(golden_dir, prec, tol):
    """End to end with the CUDA ADM network (learned-variance head: the step reads et[:, :3] in place).  fp32 mode is
    gated at the north-star tolerance; bf16 reports PSNR against the reference's final x0 (>= 30 dB, peak-to-peak 2)."""
    from nlc_b200 import svd_ddnm as SD
    from nlc_b200.unet_adm import UNetModel
    cfg = dict(weights.ADM_CONFIGS["adm_tiny"])
    cfg.pop("sigma")
    m = UNetModel(in_channels=3, precision=prec, device=dev, **{k: cfg[k] for k in KEYS}).load_state_dict(
        weights.adm_unet_state_dict(**cfg, seed=3))
    g, keys = _loop_cases(golden_dir)
    for key in keys:
        case = g[key]
        name, sy = key.split("|")
        op = _build(name, 32, 3, g["missing"], g["perm"])
        zs = iter(case["z"])
        fn = SD.ddnm_diffusion if sy == "None" else SD.ddnm_plus_diffusion
        args = () if sy == "None" else (float(sy),)
        xs, x0s = fn(case["xT"].to(dev), m, g["betas"].to(dev), g["eta"], op, case["y"].to(dev), *args,
                     config=_config(g), noise_fn=lambda like, _z=zs: next(_z).to(dev))
        if tol is not None:
            _close(xs[0], case["x_last"], tol, (key, "x_last"))
            _close(x0s[0], case["x0_last"], tol, (key, "x0_last"))
        else:
            mse = ((x0s[0] - case["x0_last"]) ** 2).mean().item()
            psnr = 10 * math.log10(4.0 / max(mse, 1e-30))
            assert psnr >= 30.0, (key, psnr)


@pytest.mark.parametrize("name", NAMES)
def test_against_oracle_seeded(name):
    """R = 64, ratio 2 (a second needle length / compression ratio), random regimes and a random fused step."""
    R, C, B = 64, 3, 2
    gen = torch.Generator().manual_seed(77)
    mask = (torch.rand(R, R, generator=gen) > 0.4).float()
    mr = torch.nonzero(mask.reshape(-1) == 0).long().reshape(-1) * 3
    missing = torch.cat([mr, mr + 1, mr + 2])
    perm = torch.randperm(R * R, generator=gen)
    op, orc = _build(name, R, C, missing, perm, ratio=2), _oracle(name, R, C, missing, perm, ratio=2)
    v, e = torch.randn(B, C * R * R, generator=gen), torch.randn(B, C * R * R, generator=gen)
    for _ in range(3):
        a, st = torch.rand((), generator=gen) * 0.9 + 0.1, torch.rand((), generator=gen) * 0.9 + 0.05
        sy, eta = float(torch.rand((), generator=gen)) * 0.5, float(torch.rand((), generator=gen))
        _close(op.Lambda(v.to(dev), a, sy, st, eta), orc.Lambda(v.clone(), a, sy, st, eta), 1e-5, (name, "Lambda"))
        _close(op.Lambda_noise(v.to(dev), a, sy, st, eta, e.to(dev)),
               orc.Lambda_noise(v.clone(), a, sy, st, eta, e.clone()), 1e-5, (name, "Lambda_noise"))
    # one fused step of each kind against the oracle loop's arithmetic (T_sampling 1000: step 500 -> 499)
    betas = torch.linspace(1e-4, 2e-2, 1000)
    xt = torch.randn(B, C, R, R, generator=gen)
    et6 = torch.randn(B, 2 * C, R, R, generator=gen)  # learned-variance layout: the first C channels are epsilon
    z = torch.randn(B, C, R, R, generator=gen)
    y = orc.A(torch.rand(B, C * R * R, generator=gen) * 2 - 1)
    at, at_next = OD.alpha_bar(betas, 500), OD.alpha_bar(betas, 499)
    for sy in (None, 0.2):
        et = et6[:, :C]
        x0 = (xt - et * (1 - at).sqrt()) / at.sqrt()
        resid = orc.A_pinv(orc.A(x0.reshape(B, -1)) - y)
        if sy is None:
            ref = at_next.sqrt() * (x0 - resid.reshape(x0.shape)) + (1 - at_next).sqrt() * 0.85 * z + \
                (1 - at_next).sqrt() * ((1 - 0.85 ** 2) ** 0.5) * et
        else:
            st = (1 - at_next).sqrt()
            ref = at_next.sqrt() * (x0 - orc.Lambda(resid, at_next.sqrt(), sy, st, 0.85).reshape(x0.shape)) + \
                orc.Lambda_noise(z.reshape(B, -1), at_next.sqrt(), sy, st, 0.85, et.reshape(B, -1)).reshape(x0.shape)
        g0, gn = op.ddnm_step(xt.to(dev), et6.to(dev), z.to(dev), y.to(dev), at, at_next, 0.85, sy)
        _close(g0, x0, 2e-6, (name, "x0", sy))
        _close(gn, ref, 2e-5, (name, "x_next", sy))


@pytest.mark.parametrize("name", NAMES)
def test_properties_at_256(name):
    R, C, B = 256, 3, 4
    gen = torch.Generator().manual_seed(9)
    mask = torch.ones(R, R)
    mask[64:192, 64:192] = 0
    mr = torch.nonzero(mask.reshape(-1) == 0).long().reshape(-1) * 3
    missing = torch.cat([mr, mr + 1, mr + 2])
    perm = torch.randperm(R * R, generator=gen)
    op = _build(name, R, C, missing, perm)
    v = torch.randn(B, C * R * R, generator=gen).to(dev)
    e = torch.randn(B, C * R * R, generator=gen).to(dev)
    # without measurement noise Lambda = V V^T = I
    assert (op.Lambda(v, 0.8, 0.0, 0.3, 0.85) - v).abs().max() < 2e-5
    # with it, Lambda only shrinks: ||Lambda v|| <= ||v||
    lv = op.Lambda(v, 0.8, 0.5, 0.1, 0.85)
    assert (lv.norm(dim=1) <= v.norm(dim=1) * (1 + 1e-5)).all()
    if name != "denoising":
        # V orthogonal, P a permutation: ||Lambda_noise(v, e)|| = ||d1 v + d2 e|| = sigma_t ||eta v + sqrt(1-eta^2) e||
        ln = op.Lambda_noise(v, 0.8, 0.0, 0.3, 0.6, e)
        want = 0.3 * (0.6 * v + 0.8 * e).norm(dim=1)
        assert ((ln.norm(dim=1) - want).abs() / want).max() < 1e-4
    # a plain DDNM step lands on the measurement: (x_next - c1 z - c2 et) / a = x0_hat with A x0_hat = y
    x_true = (torch.rand(B, C * R * R, generator=gen) * 2 - 1).to(dev)
    y = op.A(x_true)
    at, at_next, eta = 0.5, 0.6, 0.85
    x0, xn = op.ddnm_step(v, e * 0.3, e, y, at, at_next, eta, None)
    st = math.sqrt(1 - at_next)
    x0_hat = (xn.reshape(B, -1) - st * eta * e - st * math.sqrt(1 - eta ** 2) * (e * 0.3)) / math.sqrt(at_next)
    assert (op.A(x0_hat) - y).abs().max() < 2e-4
    assert torch.isfinite(x0).all()
    # DDNM+ with a tiny sigma_y behaves like the noiseless projection on the range space: A x0_hat -> y as sigma_y -> 0
    x0p, xnp = op.ddnm_step(v, e * 0.3, torch.zeros_like(e), y, at, at_next, 0.0, 1e-6)
    assert torch.isfinite(xnp).all() and (x0p - x0).abs().max() == 0
