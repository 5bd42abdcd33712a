"""GPU parity of the ADM UNet / sigma-model executors (row N1 + its G1) against the reference's golden outputs
(tests/golden/nets_adm.pt, produced by the unmodified src/unet_adm.py) and the CPU oracle.

adm_tiny exercises scale-shift norm, resblock up/down, legacy attention order with 64-channel heads and the
6-channel learned-variance output; adm_alt the other branches (emb add, conv resampling, new attention order,
fixed head count, attention at the full 32x32 resolution = 1024 tokens).
Tolerances as in tests/test_gpu_nets.py: max-norm relative 2e-3 (tf32 operands), 2e-2 (bf16 operands)."""
import os

import pytest
import torch

from oracle import adm_net, weights

pytestmark = pytest.mark.gpu
dev = torch.device("cuda:0")
TOL = {"tf32": 2e-3, "fp16": 2e-3, "bf16": 2e-2}
KEYS = ("image_size", "model_channels", "out_channels", "num_res_blocks", "attention_resolutions", "channel_mult",
        "num_heads", "num_head_channels", "use_scale_shift_norm", "resblock_updown", "use_new_attention_order")


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def _models(name, prec):
    from nlc_b200.unet_adm import SigmaModel, UNetModel
    cfg = dict(weights.ADM_CONFIGS[name])
    sg = cfg.pop("sigma")
    sd = weights.adm_unet_state_dict(**cfg, seed=3)
    ssd = weights.adm_sigma_state_dict(**sg, seed=4)
    m = UNetModel(in_channels=3, precision=prec, device=dev, **{k: cfg[k] for k in KEYS}).load_state_dict(sd)
    s = SigmaModel(dim=sg["dim"], channels=sg["channels"], n_blocks=sg["n_blocks"], num_heads=cfg["num_heads"],
                   num_head_channels=cfg["num_head_channels"], use_new_attention_order=cfg["use_new_attention_order"],
                   precision=prec, device=dev).load_state_dict(ssd)
    return cfg, sd, ssd, m, s


@pytest.mark.parametrize("prec", ["tf32", "bf16", "fp16"])
@pytest.mark.parametrize("name", ["adm_tiny", "adm_alt"])
def test_golden_reference_outputs(golden_dir, name, prec):
    _, _, _, m, s = _models(name, prec)
    g = torch.load(os.path.join(golden_dir, "nets_adm.pt"), weights_only=True)[name]
    out = m(g["x"].to(dev), g["t"].to(dev))
    feat = m.encode(g["x"].to(dev), g["t"].to(dev))
    assert out.shape == g["out"].shape and feat.shape == g["feat"].shape
    assert _rel(out.cpu(), g["out"]) < TOL[prec]
    assert _rel(feat.cpu(), g["feat"]) < TOL[prec]
    r = s(g["feat"].to(dev))
    assert r.shape == (2, 1, 1, 1)
    assert (r.cpu() - g["r"]).abs().max() < (1e-2 if prec == "bf16" else 1e-3)


@pytest.mark.parametrize("prec", ["tf32", "bf16", "fp16"])
def test_batch_of_five_vs_oracle_and_scale_folding(prec):
    """Odd batch (ragged last M tile), forward_and_encode, and the folded input scale."""
    cfg, sd, ssd, m, s = _models("adm_tiny", prec)
    R = cfg["image_size"]
    g = torch.Generator().manual_seed(8)
    x = torch.randn(5, 3, R, R, generator=g)
    t = torch.tensor([999.0, 250.0, 3.0, 0.0, 512.0])
    sc = torch.rand(5, generator=g) + 0.2
    with torch.no_grad():
        ref, feat = adm_net.unet_forward(sd, x * sc.view(-1, 1, 1, 1), t, cfg, return_feat=True)
        r_ref = adm_net.sigma_forward(ssd, feat, cfg)
    out = m.forward_scaled(x.to(dev), t.to(dev), sc.to(dev)).clone()
    f = m.encode_scaled(x.to(dev), t.to(dev), sc.to(dev)).clone().permute(0, 3, 1, 2)
    assert _rel(out.cpu(), ref) < TOL[prec]
    assert _rel(f.cpu(), feat) < TOL[prec]
    assert (s(f).cpu() - r_ref).abs().max() < (2e-2 if prec == "bf16" else 2e-3)
