"""Where one NLC timestep goes: per-op CUDA-event times (ops.STATS.op_timer) for one encode + sigma-model +
forward of a network family at a given batch.  Usage:
    python scripts/step_profile.py {c2|adm256|adm_tiny} BATCH {bf16|tf32} [reps]
Event bracketing serialises nothing (single stream) but adds ~2 us per launch; totals are also reported from one
un-instrumented pass."""
import collections
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nlc_b200 import ops
from oracle import weights  # seeded synthetic state_dicts only

name, B, prec = sys.argv[1], int(sys.argv[2]), sys.argv[3]
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 3
dev = torch.device("cuda:0")
if name.startswith("adm"):
    from nlc_b200.unet_adm import SigmaModel, UNetModel
    cfg = dict(weights.ADM_CONFIGS[name])
    sg = cfg.pop("sigma")
    keys = ("image_size", "model_channels", "out_channels", "num_res_blocks", "attention_resolutions", "channel_mult",
            "num_heads", "num_head_channels", "use_scale_shift_norm", "resblock_updown", "use_new_attention_order")
    m = UNetModel(in_channels=3, precision=prec, device=dev, **{k: cfg[k] for k in keys}).load_state_dict(
        weights.adm_unet_state_dict(**cfg, seed=3))
    s = SigmaModel(dim=sg["dim"], channels=sg["channels"], n_blocks=sg["n_blocks"], num_heads=cfg["num_heads"],
                   num_head_channels=cfg["num_head_channels"], precision=prec, device=dev).load_state_dict(
        weights.adm_sigma_state_dict(**sg, seed=4))
    R = cfg["image_size"]
else:
    from nlc_b200.unet_ddim import SigmaModel, UNetModel
    cfg = weights.CONFIGS[name]
    m = UNetModel(**cfg["unet"], precision=prec, device=dev).load_state_dict(
        weights.ddim_unet_state_dict(**cfg["unet"], seed=3))
    s = SigmaModel(**cfg["sigma"], precision=prec, device=dev).load_state_dict(
        weights.ddim_sigma_state_dict(**cfg["sigma"], seed=4))
    R = cfg["unet"]["image_size"]
x = torch.randn(B, 3, R, R, device=dev)
t = torch.full((B,), 500.0, device=dev)
sc = torch.full((B,), 0.3, device=dev)


def step():
    f = m.encode_scaled(x, t, sc)
    s.forward_nhwc(f)
    m.forward_scaled(x, t, sc)


for _ in range(2):
    step()
torch.cuda.synchronize()
print("device memory: %.2f GB (engine buffers %.2f GB)" % (torch.cuda.memory_allocated() / 1e9,
                                                          (m.eng.bytes_allocated() + s.eng.bytes_allocated()) / 1e9))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print("%s B=%d %s: %.3f ms per NLC timestep (encode + sigma + forward), %.2f ms/img" % (name, B, prec, ms, ms / B))

ops.STATS.op_timer = []
for _ in range(reps):
    step()
torch.cuda.synchronize()
rec, ops.STATS.op_timer = ops.STATS.op_timer, None
agg = collections.OrderedDict()
for nm, a, b, fl in rec:
    d = agg.setdefault(nm, [0, 0.0, 0.0])
    d[0] += 1
    d[1] += a.elapsed_time(b)
    d[2] += fl
tot = sum(v[1] for v in agg.values())
print("instrumented total %.3f ms / timestep" % (tot / reps))
conv_ms = sum(v[1] for k, v in agg.items() if k.startswith("conv_tc"))
conv_fl = sum(v[2] for k, v in agg.items() if k.startswith("conv_tc"))
print("conv_tc: %.1f%% of step, %.1f TFLOP/s aggregate" % (100 * conv_ms / tot, conv_fl / conv_ms / 1e9))
for nm, (n, msum, fl) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    line = "%-44s n=%4d  %9.3f ms  %5.1f%%" % (nm, n // reps, msum / reps, 100 * msum / tot)
    if fl:
        line += "  %7.1f TFLOP/s" % (fl / msum / 1e9)
    print(line)
