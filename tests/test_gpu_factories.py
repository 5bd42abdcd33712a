"""the reference's model factories (src/script_util.py) and checkpoint formats drive the CUDA networks.
The factories get the reference's own keyword arguments / YAML-shaped config, the weights arrive through the checkpoint
readers (a DDIM `.ckpt` list with EMA for the DDIM UNet, a plain `.pt` for ADM), and the outputs are compared with the
unmodified reference's golden outputs for the same weights (tests/golden/nets_tiny.pt, nets_adm.pt)."""
import os
import types

import pytest
import torch

from oracle import weights

pytestmark = pytest.mark.gpu
dev = torch.device("cuda:0")


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def test_adm_factory_and_pt_checkpoint(golden_dir, tmp_path):
    from nlc_b200 import checkpoints as CK, script_util as SU
    cfg = dict(weights.ADM_CONFIGS["adm_tiny"])
    sg = cfg.pop("sigma")
    path = tmp_path / "adm.pt"
    torch.save(weights.adm_unet_state_dict(**cfg, seed=3), path)
    eps, sig, feat_shape = SU.create_sigma_eps_model(
        image_size=32, num_channels=128, num_res_blocks=1, channel_mult="1,2", learn_sigma=True,
        attention_resolutions="16", num_heads=4, num_head_channels=64, use_scale_shift_norm=True, resblock_updown=True,
        use_new_attention_order=False, sigma_block=2, precision="tf32", device=dev)
    assert feat_shape == (256, 16, 16) == (sg["channels"], sg["dim"], sg["dim"])
    CK.load_eps_model(eps, str(path))
    sig.load_state_dict(weights.adm_sigma_state_dict(**sg, seed=4))
    g = torch.load(os.path.join(golden_dir, "nets_adm.pt"), weights_only=True)["adm_tiny"]
    out = eps(g["x"].to(dev), g["t"].to(dev))
    assert out.shape == g["out"].shape and _rel(out.cpu(), g["out"]) < 2e-3
    assert (sig(g["feat"].to(dev)).cpu() - g["r"]).abs().max() < 5e-4
    with pytest.raises(ValueError):
        SU.create_sigma_eps_model(image_size=48, num_channels=64, num_res_blocks=1, device=dev)


def test_ddim_factory_and_ema_ckpt(golden_dir, tmp_path):
    from nlc_b200 import checkpoints as CK, script_util as SU
    cfg = weights.CONFIGS["tiny"]
    sd = weights.ddim_unet_state_dict(**cfg["unet"], seed=3)
    # DDIM checkpoint layout: DataParallel keys, stale raw weights in slot 0, the weights that count in the EMA slot
    raw = {"module." + k: torch.zeros_like(v) for k, v in sd.items()}
    path = tmp_path / "model.ckpt"
    torch.save([raw, {}, 0, 0, dict(sd)], path)
    config = types.SimpleNamespace(
        model=types.SimpleNamespace(ch=128, out_ch=3, ch_mult=[1, 2], num_res_blocks=1, attn_resolutions=[8], dropout=0.0,
                                    in_channels=3, resamp_with_conv=True, sigma_block=2, sigma_dropout=0.0, feat_layer=0,
                                    type="simple"),
        data=types.SimpleNamespace(image_size=16), diffusion=types.SimpleNamespace(num_diffusion_timesteps=1000))
    eps, sig, feat_shape = SU.create_simple_sigma_eps_model(config, precision="tf32", device=dev)
    assert feat_shape == (256, 8, 8)
    CK.load_eps_model(eps, str(path))
    sig.load_state_dict(weights.ddim_sigma_state_dict(**cfg["sigma"], seed=4))
    g = torch.load(os.path.join(golden_dir, "nets_tiny.pt"), weights_only=True)
    assert _rel(eps(g["x"].to(dev), g["t"].to(dev)).cpu(), g["out"]) < 2e-3
    assert (sig(g["feat"].to(dev)).cpu() - g["r"]).abs().max() < 5e-4


def test_edm_factory(golden_dir):
    from nlc_b200 import script_util as SU
    cfg = dict(weights.EDM_CONFIGS["edm_tiny"])
    sg = cfg.pop("sigma")
    eps, sig, feat_shape = SU.create_edm_sigma_eps_model(**cfg, sigma_block=sg["n_blocks"], precision="tf32", device=dev)
    assert feat_shape == (sg["channels"], sg["dim"], sg["dim"])
    eps.load_state_dict(weights.edm_unet_state_dict(**cfg, seed=3))
    sig.load_state_dict(weights.edm_sigma_state_dict(**sg, seed=4))
    g = torch.load(os.path.join(golden_dir, "nets_edm.pt"), weights_only=True)
    assert _rel(eps(g["x"].to(dev), g["c_noise"].to(dev)).cpu(), g["out"]) < 2e-3
    assert (sig(g["feat"].to(dev)).cpu() - g["r"]).abs().max() < 1e-3
