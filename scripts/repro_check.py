"""Bit-reproducibility of the ADM-256 forward (same plan, same launches, twice) under the kernel switches of the context:
    python scripts/repro_check.py [precision] [reps]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nlc_b200 import _lib
from nlc_b200 import synthetic_weights as weights
from nlc_b200.unet_adm import UNetModel

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 4
dev = torch.device("cuda:0")
KEYS = ("image_size", "model_channels", "out_channels", "num_res_blocks", "attention_resolutions", "channel_mult",
        "num_heads", "num_head_channels", "use_scale_shift_norm", "resblock_updown", "use_new_attention_order")
cfg = dict(weights.ADM_CONFIGS["adm256"])
cfg.pop("sigma")
m = UNetModel(in_channels=3, precision=prec, device=dev, **{k: cfg[k] for k in KEYS}).load_state_dict(
    weights.adm_unet_state_dict(**cfg, seed=3))
g = torch.Generator().manual_seed(9)
x = torch.randn(3, 3, 256, 256, generator=g).to(dev)
t = torch.tensor([900.0, 400.0, 20.0], device=dev)
sc = torch.tensor([0.05, 0.3, 0.9], device=dev)
ctx = _lib.ctx(0)
for tma in ((1,) if os.environ.get("NLC_TMA_EPI_MASK") else (0, 1, 2)):
    for att in ((1,) if os.environ.get("NLC_TMA_EPI_MASK") else (0, 1)):
        _lib.check(_lib.lib().nlc_ctx_set(ctx, b"tma_epi", tma))
        _lib.check(_lib.lib().nlc_ctx_set(ctx, b"attn_onepass", att))
        ref = m.forward_scaled(x, t, sc).clone()
        worst, bad = 0.0, 0
        for _ in range(reps):
            again = m.forward_scaled(x, t, sc).clone()
            d = (again - ref).abs()
            worst = max(worst, d.max().item())
            bad = max(bad, int((d > 0).sum().item()))
        print("%s tma_epi=%d attn_onepass=%d: max |diff| over %d repeats %.3e, differing elements <= %d of %d" % (
            prec, tma, att, reps, worst, bad, ref.numel()), flush=True)
