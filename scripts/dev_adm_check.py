"""Development check: ADM mirror vs the CPU oracle, block by block (first diverging block), then the adm256 plan
at a small batch with per-step synchronisation (NLC_SYNC=1) to localise kernel faults."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nlc_b200.unet_adm import SigmaModel, UNetModel
from oracle import adm_net, weights

dev = torch.device("cuda:0")
KEYS = ("image_size", "model_channels", "out_channels", "num_res_blocks", "attention_resolutions", "channel_mult",
        "num_heads", "num_head_channels", "use_scale_shift_norm", "resblock_updown", "use_new_attention_order")


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def blockwise(name, prec):
    cfg = dict(weights.ADM_CONFIGS[name])
    sg = cfg.pop("sigma")
    sd = weights.adm_unet_state_dict(**cfg, seed=3)
    m = UNetModel(in_channels=3, precision=prec, device=dev, **{k: cfg[k] for k in KEYS}).load_state_dict(sd)
    g = torch.Generator().manual_seed(8)
    R = cfg["image_size"]
    x = torch.randn(2, 3, R, R, generator=g)
    t = torch.tensor([999.0, 250.0])
    cp = adm_net._prepare(cfg)
    with torch.no_grad():
        emb = adm_net._embed(sd, t, cp)
        h, hs = x, []
        for i in range(adm_net._n_blocks(sd, "input_blocks")):
            h = adm_net._run_block(sd, "input_blocks.%d." % i, h, emb, cp, False)
            hs.append(h)
        mid = adm_net._run_block(sd, "middle_block.", h, emb, cp, False)
        outs, h = [], mid
        skips = list(hs)
        for i in range(adm_net._n_blocks(sd, "output_blocks")):
            h = adm_net._run_block(sd, "output_blocks.%d." % i, torch.cat([h, skips.pop()], dim=1), emb, cp, True)
            outs.append(h)
        ref = adm_net.unet_forward(sd, x, t, cfg)
    out = m.forward_scaled(x.to(dev), t.to(dev)).clone()
    P = m._plan(2)
    print("== %s %s: final rel %.3e" % (name, prec, rel(out.cpu(), ref)))
    e = (P["emb"].cpu())
    print("   temb rel %.3e" % rel(P["temb"].cpu(), emb))
    cat = P["cat"]
    n = len(hs)
    for k in range(n):
        c32, _, c1 = cat[k]
        mine = c32[..., c1:c1 + m.skip_ch[k]].permute(0, 3, 1, 2).cpu()
        print("   skip %2d %-18s rel %.3e" % (k, tuple(hs[k].shape), rel(mine, hs[k])))
    c32, _, c1 = cat[n - 1]
    print("   middle rel %.3e" % rel(c32[..., :c1].permute(0, 3, 1, 2).cpu(), mid))
    for j in range(n - 1):
        k = n - 2 - j
        c32, _, c1 = cat[k]
        print("   out %2d %-18s rel %.3e" % (j, tuple(outs[j].shape), rel(c32[..., :c1].permute(0, 3, 1, 2).cpu(), outs[j])))


for name in ("adm_tiny", "adm_alt"):
    for prec in ("tf32",):
        try:
            blockwise(name, prec)
        except Exception as ex:
            print("blockwise %s %s failed: %r" % (name, prec, ex))

# ---- adm256 small batch with per-step sync
os.environ["NLC_SYNC"] = "1"
cfg = dict(weights.ADM_CONFIGS["adm256"])
sg = cfg.pop("sigma")
m = UNetModel(in_channels=3, precision="bf16", device=dev, **{k: cfg[k] for k in KEYS}).load_state_dict(
    weights.adm_unet_state_dict(**cfg, seed=3))
s = SigmaModel(dim=sg["dim"], channels=sg["channels"], n_blocks=sg["n_blocks"], num_heads=cfg["num_heads"],
               num_head_channels=cfg["num_head_channels"], precision="bf16", device=dev).load_state_dict(
    weights.adm_sigma_state_dict(**sg, seed=4))
for B in (8, 2):
    x = torch.randn(B, 3, 256, 256, device=dev)
    t = torch.full((B,), 500.0, device=dev)
    try:
        f = m.encode_scaled(x, t, None)
        print("adm256 B=%d encode ok" % B, float(f.abs().mean()))
        r = s.forward_nhwc(f)
        print("adm256 B=%d sigma ok" % B, r.flatten()[:4].tolist())
        o = m.forward_scaled(x, t, None)
        print("adm256 B=%d forward ok" % B, float(o.abs().mean()))
    except Exception as ex:
        print("adm256 B=%d: %s" % (B, str(ex)[:600]))
        break
