"""Development check: first ADM ResBlock (scale-shift norm) stage by stage against torch."""
import os
import sys

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nlc_b200.unet_adm import UNetModel
from oracle import adm_net, weights

dev = torch.device("cuda:0")
KEYS = ("image_size", "model_channels", "out_channels", "num_res_blocks", "attention_resolutions", "channel_mult",
        "num_heads", "num_head_channels", "use_scale_shift_norm", "resblock_updown", "use_new_attention_order")
cfg = dict(weights.ADM_CONFIGS["adm_tiny"])
cfg.pop("sigma")
sd = weights.adm_unet_state_dict(**cfg, seed=3)
m = UNetModel(in_channels=3, precision="tf32", device=dev, **{k: cfg[k] for k in KEYS}).load_state_dict(sd)
g = torch.Generator().manual_seed(8)
x = torch.randn(2, 3, 32, 32, generator=g)
t = torch.tensor([999.0, 250.0])
P = m._plan(2)
m._stage(P, x.to(dev), t.to(dev), None)
P["emb_n"][0] = m.emb_total


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def scratch(tag, shape, dtype=torch.float32):
    n = 1
    for s_ in shape:
        n *= s_
    return m.eng._scratch[(tag, dtype)][:n].view(*shape).float().permute(0, 3, 1, 2).cpu()


with torch.no_grad():
    emb = adm_net._embed(sd, t, cfg)
    h0 = F.conv2d(x, sd["input_blocks.0.0.weight"], sd["input_blocks.0.0.bias"], padding=1)
    p = "input_blocks.1.0."
    a1 = F.silu(F.group_norm(h0, 32, sd[p + "in_layers.0.weight"], sd[p + "in_layers.0.bias"], eps=1e-5))
    h = F.conv2d(a1, sd[p + "in_layers.2.weight"], sd[p + "in_layers.2.bias"], padding=1)
    e = F.linear(F.silu(emb), sd[p + "emb_layers.1.weight"], sd[p + "emb_layers.1.bias"])
    sc, sh = e.chunk(2, dim=1)
    a2 = F.silu(F.group_norm(h, 32, sd[p + "out_layers.0.weight"], sd[p + "out_layers.0.bias"], eps=1e-5)
                * (1 + sc[:, :, None, None]) + sh[:, :, None, None])
    o = F.conv2d(a2, sd[p + "out_layers.3.weight"], sd[p + "out_layers.3.bias"], padding=1) + h0
steps = P["enc"]
for i in range(5):
    steps[i]()
torch.cuda.synchronize()
w = m.input_blocks[0][0][1]
print("emb_off", w.emb_off, "cout", w.cout, "emb_total", m.emb_total, "scale_shift", w.scale_shift)
print("emb rows rel", rel(P["emb"][:, w.emb_off:w.emb_off + 2 * w.cout].cpu(), e))
for i, (tag, ref) in zip(range(5, 9), (("rb.a1", a1), ("rb.h", h), ("rb.a2", a2), (None, o))):
    print("step", i, getattr(steps[i], "label", "?"))
    steps[i]()
    torch.cuda.synchronize()
    if tag:
        print("   %s rel %.3e (|ref| max %.3f)" % (tag, rel(scratch(tag, (2, 32, 32, 128)), ref), ref.abs().max()))
    else:
        c32, _, c1 = P["cat"][1]
        mine = c32[..., c1:c1 + 128].permute(0, 3, 1, 2).cpu()
        print("   out rel %.3e  |mine| %.4f |ref| %.4f" % (rel(mine, o), mine.abs().mean(), o.abs().mean()))
