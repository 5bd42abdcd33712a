"""The accuracy mode (precision="fp32", NLC_F32X3): unrounded fp32 operands, every tensor-core product computed as
three kind::tf32 MMAs on an in-kernel hi/lo split (csrc/conv_tc.cu MODE 2).

This is the mode in which the north-star's floating-point tolerance is gated: **every timestep's corrected sigma,
eps and x_{t-1} within 1e-4 relative of the reference PyTorch sampler** (teacher-forced on the reference's own x_t,
tests/golden/*.pt, produced by the unmodified reference on CPU in fp32).  Network outputs are held to 1e-4 of the
output's max magnitude as well; the single kernels to fp32 summation-order level against float64 torch.
"""
import math
import os

import pytest
import torch
import torch.nn.functional as F

from oracle import weights

pytestmark = pytest.mark.gpu
dev = torch.device("cuda:0")
NORTH_STAR_TOL = 1e-4


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def _l2rel(a, b):
    return (torch.linalg.vector_norm(a.double() - b.double()) /
            torch.linalg.vector_norm(b.double()).clamp_min(1e-30)).item()


# ------------------------------------------------------------------------------------------------ kernels
CONV_CASES = [
    # B, H, W, Cin, Cout, stride, pad(l,r,t,b), k
    (4, 16, 16, 128, 128, 1, (1, 1, 1, 1), 3),
    (5, 4, 4, 256, 512, 1, (1, 1, 1, 1), 3),
    (3, 32, 32, 256, 768, 1, (0, 0, 0, 0), 1),
    (4, 32, 32, 128, 128, 2, (0, 1, 0, 1), 3),
    (1, 256, 256, 64, 64, 1, (1, 1, 1, 1), 3),
    (2, 64, 64, 256, 256, 1, (1, 1, 1, 1), 3),     # many tiles per CTA: the 3/4-stage ring wraps several times
    (130, 1, 1, 256, 128, 1, (0, 0, 0, 0), 1),
]


@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_tc_fp32_mode_vs_float64(case):
    """Plain (unrounded) fp32 data.  One tf32 MMA would be off by ~3e-4 of max|y|; the split products must land at
    fp32 level: 2e-6 of max|y| against the float64 convolution of the same data."""
    from nlc_b200 import ops
    from nlc_b200._lib import NLC_F32X3
    B, H, W, Cin, Cout, stride, pad, k = case
    g = torch.Generator().manual_seed(11)
    x = torch.randn(B, Cin, H, W, generator=g).to(dev)
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).to(dev)
    b = torch.randn(Cout, generator=g).to(dev)
    resid = torch.randn(B, H // stride, W // stride, Cout, generator=g).to(dev)
    ref = F.conv2d(F.pad(x.double(), pad), w.double(), b.double(), stride=stride) + resid.double().permute(0, 3, 1, 2)
    Ho, Wo = ref.shape[2:]
    xa = ops.Act(x.permute(0, 2, 3, 1).contiguous())
    segs = [(0, kh - pad[2], kw - pad[0], 0, Cin) for kh in range(k) for kw in range(k)]
    o32 = ops.Act(torch.full((B, Ho, Wo, Cout), float("nan"), device=dev))
    oop = ops.Act(torch.zeros(B, Ho, Wo, Cout, device=dev))
    ops.conv_tc([xa], segs, ops.pack_conv_weight(w, NLC_F32X3), Cout, B, Ho, Wo, NLC_F32X3, stride=stride, bias=b,
                resid=ops.Act(resid), out_f32=o32, out_op=oop)
    torch.cuda.synchronize()
    err = _rel(o32.t.permute(0, 3, 1, 2).double(), ref)
    assert err < 2e-6, err
    assert torch.equal(oop.t, o32.t)  # the operand copy is the unrounded fp32 value


@pytest.mark.parametrize("case", [(3, 16, 1, 512, False), (2, 64, 4, 64, True), (3, 256, 1, 256, False),
                                  (2, 1024, 4, 64, False), (2, 256, 4, 64, True)])
def test_attention_fp32_mode(case):
    """Both attention GEMMs take an activation as their right-hand operand, so the hi/lo split of the B tile happens
    in the kernel too.  5e-6 of max|out| against float64 attention."""
    from nlc_b200 import ops
    from nlc_b200._lib import NLC_F32X3
    B, T, heads, dh, legacy = case
    C = heads * dh
    g = torch.Generator().manual_seed(5)
    qkv = torch.randn(B, T, 3 * C, generator=g).to(dev)
    f = qkv.double()
    if legacy:
        v5 = f.view(B, T, heads, 3, dh)
        q, k, v = v5[:, :, :, 0], v5[:, :, :, 1], v5[:, :, :, 2]
        offs = (0, dh, 2 * dh, 3 * dh)
    else:
        q, k, v = [f[:, :, i * C:(i + 1) * C].view(B, T, heads, dh) for i in range(3)]
        offs = (0, C, 2 * C, dh)
    scale = dh ** -0.5
    w = torch.softmax(torch.einsum("bthd,bshd->bhts", q, k) * scale, dim=-1)
    ref = torch.einsum("bhts,bshd->bthd", w, v).reshape(B, T, C)
    side = int(T ** 0.5)
    out = ops.Act(torch.zeros(B, side, T // side, C, device=dev))
    ws = torch.zeros(max(ops.attention_ws(NLC_F32X3, B, T, heads, dh), 16), device=dev, dtype=torch.uint8)
    ops.attention(ops.Act(qkv.view(B, side, T // side, 3 * C)), NLC_F32X3, offs[0], offs[1], offs[2], offs[3], heads,
                  dh, scale, out, ws)
    err = _rel(out.t.view(B, T, C).double(), ref)
    assert err < 5e-6, err


# ------------------------------------------------------------------------------------------------ networks
def test_ddim_unet_and_sigma_model_golden(golden_dir):
    from nlc_b200.unet_ddim import SigmaModel, UNetModel
    cfg = weights.CONFIGS["tiny"]
    m = UNetModel(**cfg["unet"], precision="fp32", device=dev).load_state_dict(
        weights.ddim_unet_state_dict(**cfg["unet"], seed=3))
    s = SigmaModel(**cfg["sigma"], precision="fp32", device=dev).load_state_dict(
        weights.ddim_sigma_state_dict(**cfg["sigma"], seed=4))
    g = torch.load(os.path.join(golden_dir, "nets_tiny.pt"), weights_only=True)
    out = m(g["x"].to(dev), g["t"].to(dev))
    feat = m.encode(g["x"].to(dev), g["t"].to(dev))
    assert _rel(out.cpu(), g["out"]) < NORTH_STAR_TOL
    assert _rel(feat.cpu(), g["feat"]) < NORTH_STAR_TOL
    assert (s(g["feat"].to(dev)).cpu() - g["r"]).abs().max() < NORTH_STAR_TOL


@pytest.mark.parametrize("name,B", [("c1", 3), ("c2", 2)])
def test_benchmark_architectures_vs_oracle(name, B):
    """The c1 / c2 networks of BASELINE.json (K up to 9216 per convolution) against the CPU oracle."""
    from nlc_b200.unet_ddim import SigmaModel, UNetModel
    from oracle import ddim_net
    cfg = weights.CONFIGS[name]
    sd = weights.ddim_unet_state_dict(**cfg["unet"], seed=3)
    ssd = weights.ddim_sigma_state_dict(**cfg["sigma"], seed=4)
    m = UNetModel(**cfg["unet"], precision="fp32", device=dev).load_state_dict(sd)
    s = SigmaModel(**cfg["sigma"], precision="fp32", device=dev).load_state_dict(ssd)
    R = cfg["unet"]["image_size"]
    g = torch.Generator().manual_seed(5)
    x = torch.randn(B, 3, R, R, generator=g)
    t = torch.tensor([999.0, 250.0, 3.0][:B])
    with torch.no_grad():
        ref, feat = ddim_net.unet_forward(sd, x, t, return_feat=True)
        r_ref = ddim_net.sigma_forward(ssd, feat)
    out, f = m.forward_and_encode(x.to(dev), t.to(dev))
    e_out, e_feat, e_r = _rel(out.cpu(), ref), _rel(f.cpu(), feat), (s(f).cpu() - r_ref).abs().max().item()
    print("fp32 mode, %s: out %.2e feat %.2e r %.2e" % (name, e_out, e_feat, e_r))
    assert e_out < NORTH_STAR_TOL and e_feat < NORTH_STAR_TOL and e_r < NORTH_STAR_TOL


@pytest.mark.parametrize("name", ["adm_tiny", "adm_alt"])
def test_adm_unet_and_sigma_model_golden(golden_dir, name):
    from nlc_b200.unet_adm import SigmaModel, UNetModel
    keys = ("image_size", "model_channels", "out_channels", "num_res_blocks", "attention_resolutions", "channel_mult",
            "num_heads", "num_head_channels", "use_scale_shift_norm", "resblock_updown", "use_new_attention_order")
    cfg = dict(weights.ADM_CONFIGS[name])
    sg = cfg.pop("sigma")
    m = UNetModel(in_channels=3, precision="fp32", device=dev, **{k: cfg[k] for k in keys}).load_state_dict(
        weights.adm_unet_state_dict(**cfg, seed=3))
    s = SigmaModel(dim=sg["dim"], channels=sg["channels"], n_blocks=sg["n_blocks"], num_heads=cfg["num_heads"],
                   num_head_channels=cfg["num_head_channels"], use_new_attention_order=cfg["use_new_attention_order"],
                   precision="fp32", device=dev).load_state_dict(weights.adm_sigma_state_dict(**sg, seed=4))
    g = torch.load(os.path.join(golden_dir, "nets_adm.pt"), weights_only=True)[name]
    out = m(g["x"].to(dev), g["t"].to(dev))
    feat = m.encode(g["x"].to(dev), g["t"].to(dev))
    assert _rel(out.cpu(), g["out"]) < NORTH_STAR_TOL
    assert _rel(feat.cpu(), g["feat"]) < NORTH_STAR_TOL
    assert (s(g["feat"].to(dev)).cpu() - g["r"]).abs().max() < NORTH_STAR_TOL


def test_edm_unet_and_sigma_model_golden(golden_dir):
    from nlc_b200.edm_networks import SigmaModel, SongUNet
    cfg = dict(weights.EDM_CONFIGS["edm_tiny"])
    sg = cfg.pop("sigma")
    m = SongUNet(precision="fp32", device=dev, **cfg).load_state_dict(weights.edm_unet_state_dict(**cfg, seed=3))
    s = SigmaModel(dim=sg["dim"], channels=sg["channels"], n_blocks=sg["n_blocks"], precision="fp32",
                   device=dev).load_state_dict(weights.edm_sigma_state_dict(**sg, seed=4))
    g = torch.load(os.path.join(golden_dir, "nets_edm.pt"), weights_only=True)
    out = m(g["x"].to(dev), g["c_noise"].to(dev))
    feat = m.encode(g["x"].to(dev), g["c_noise"].to(dev))
    assert _rel(out.cpu(), g["out"]) < NORTH_STAR_TOL
    assert _rel(feat.cpu(), g["feat"]) < NORTH_STAR_TOL
    assert (s(g["feat"].to(dev)).cpu() - g["r"]).abs().max() < NORTH_STAR_TOL


# ------------------------------------------------------------------------------------------------ the sampling step
def _experiment(kind, eta, var, n_steps=6, start=20.0):
    from nlc_b200.experiments import ImageExperiment
    from nlc_b200.schedulers import get_sampler
    from nlc_b200.unet_ddim import SigmaModel, UNetModel
    cfg = weights.CONFIGS["tiny"]
    R = cfg["unet"]["image_size"]
    m = UNetModel(**cfg["unet"], precision="fp32", device=dev).load_state_dict(
        weights.ddim_unet_state_dict(**cfg["unet"], seed=3))
    s = SigmaModel(**cfg["sigma"], precision="fp32", device=dev).load_state_dict(
        weights.ddim_sigma_state_dict(**cfg["sigma"], seed=4))
    sch = get_sampler(kind, 1000, n_steps, start_sigma=start, sampler_var=var, eta=eta).to(dev)
    exp = ImageExperiment(m, sch, batch_size=2, data_shape=(3, R, R), seed=5, device=dev)
    exp.set_model(m, s, learn_epsvar=False)
    exp.set_norm_maxmin(0.0, 30.0)
    exp.set_clip_fn("clamp")
    return exp, sch


@pytest.fixture(scope="module")
def golden_loops(golden_dir):
    return torch.load(os.path.join(golden_dir, "denoise_loop_tiny.pt"), weights_only=True)


@pytest.mark.parametrize("key", ["ddim|0.0|none", "ddim_simple_orig|0.85|none", "ddim|0.5|fixedsmall",
                                 "ddpm|1.0|fixedlarge", "ddpm_orig|1.0|fixedsmall", "ddim_orig|0.3|fixedlarge",
                                 "ddim_simple|0.2|none", "ddim_simple_drag|0.2|none"])
def test_every_timestep_within_1e4_of_the_reference(golden_loops, key):
    """North-star gate (fp32 mode): sigma_hat, sigma_hat_prev, eps, x0_hat and x_{t-1} of every timestep, computed
    from the reference's own x_t, within 1e-4 L2-relative of the reference's dump of that step."""
    kind, eta, var = key.split("|")
    case = golden_loops[key]
    exp, sch = _experiment(kind, float(eta), var)
    worst = 0.0
    for i in range(len(case["eps"])):
        xt = case["xt"][i].to(dev)
        eps, lv, s_t, s_p = exp.get_denoise_vector(xt, int(case["timesteps"][i]), sch.sampling_sigmas[i:i + 1],
                                                   sch.sampling_sigmas[i + 1:i + 2], "pred", True, True)
        x0h = sch.pred_xstart(xt, eps, s_t, clip=exp.clip_mode)
        noise = case["noises"][i].to(dev) if case["noises"] else None
        xp = sch.pred_xprev(x0=x0h, eps=eps, sigma_t=s_t, sigma_prev=s_p, xt=xt, log_variance=lv, noise=noise)
        errs = dict(sigma=_l2rel(s_t.reshape(-1).cpu(), case["sigma_t"][i]),
                    sigma_prev=_l2rel(s_p.reshape(-1).cpu(), case["sigma_prev"][i]),
                    eps=_l2rel(eps.cpu(), case["eps"][i]), x0=_l2rel(x0h.cpu(), case["x0_hat"][i]),
                    x_prev=_l2rel(xp.cpu(), case["x_prev"][i]))
        for name, e in errs.items():
            assert e < NORTH_STAR_TOL, (key, i, name, e)
        worst = max(worst, max(errs.values()))
    print("fp32 mode, %s: worst per-step relative error %.2e" % (key, worst))


def test_free_running_trajectory(golden_loops):
    """The whole 6-step loop without teacher forcing: final image >= 80 dB PSNR against the reference (peak-to-peak 2);
    the bf16 throughput mode reaches 38-43 dB on the same case (tests/test_gpu_sampler.py)."""
    case = golden_loops["ddim_simple_orig|0.85|none"]
    exp, sch = _experiment("ddim_simple_orig", 0.85, "none")
    xT = (case["z"] / (1 / (case["sigmas"][0] ** 2 + 1)).sqrt()).to(dev)
    out, _ = exp.denoise_loop(shape=tuple(xT.shape), xT=xT, style="pred", norm_eps=True, refine_prior_sigma=True,
                              return_log=False, noise_fn=lambda i, like: case["noises"][i].to(dev))
    mse = torch.mean((out - case["final"]) ** 2).item()
    psnr = 10 * math.log10(4.0 / max(mse, 1e-30))
    assert psnr >= 80.0, psnr


def test_edm_denoise_vector_within_1e4(golden_dir):
    """EDM path (rows D2/L3): every get_denoise_vector call of the reference's Heun sampler, teacher-forced."""
    import test_gpu_edm as E
    exp = E._experiment("fp32")
    golden = torch.load(os.path.join(golden_dir, "edm_sampler_tiny.pt"), weights_only=False)
    for key in E.CASES:
        style, ne, refine, _ = key.split("|")
        for c in golden[key]["calls"]:
            args = [v.reshape(()) if v.numel() == 1 else v.view(-1, 1, 1, 1).to(dev)
                    for v in (c["sigma_in"], c["sigma_prev_in"])]
            eps, _, s_t, _ = exp.get_denoise_vector(c["xt"].to(dev), args[0], args[1], style=c["style"],
                                                    norm_eps=bool(int(ne[0])), refine_prior_sigma=bool(int(refine)))
            assert _l2rel(eps.cpu(), c["eps"]) < NORTH_STAR_TOL, (key, c["style"])
            mine, ref = s_t.reshape(-1).cpu().double(), c["sigma_t"]
            assert _l2rel(mine.expand(2), ref.expand(2)) < NORTH_STAR_TOL, (key, c["style"])
