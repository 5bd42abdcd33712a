"""PyTorch-eager baseline on the same GPU: the reference's networks (functional restatement in oracle/, pinned bit-exact to
the reference on CPU) evaluated by torch / cuDNN for one NLC timestep (UNet encode + sigma-model + UNet forward), the
way the reference itself runs on a GPU: fp32 tensors, cuDNN convolutions with TF32 allowed (torch's default), fp32
matmuls.  This is the "kernel to beat" figure of SURVEY section 8(d); it is a reported baseline, not a product path.

    python scripts/torch_eager_baseline.py {c2|adm256} BATCH [reps]
"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import weights  # noqa: E402

name, B = sys.argv[1], int(sys.argv[2])
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda:0")
torch.backends.cudnn.benchmark = True
if name.startswith("adm"):
    from oracle import adm_net
    cfg = dict(weights.ADM_CONFIGS[name])
    sg = cfg.pop("sigma")
    sd = {k: v.to(dev) for k, v in weights.adm_unet_state_dict(**cfg, seed=3).items()}
    ssd = {k: v.to(dev) for k, v in weights.adm_sigma_state_dict(**sg, seed=4).items()}
    R = cfg["image_size"]
    enc = lambda x, t: adm_net.unet_encode(sd, x, t, cfg)
    fwd = lambda x, t: adm_net.unet_forward(sd, x, t, cfg)
    sig = lambda f: adm_net.sigma_forward(ssd, f, cfg)
else:
    from oracle import ddim_net
    cfg = weights.CONFIGS[name]
    sd = {k: v.to(dev) for k, v in weights.ddim_unet_state_dict(**cfg["unet"], seed=3).items()}
    ssd = {k: v.to(dev) for k, v in weights.ddim_sigma_state_dict(**cfg["sigma"], seed=4).items()}
    R = cfg["unet"]["image_size"]
    enc = lambda x, t: ddim_net.unet_encode(sd, x, t)
    fwd = lambda x, t: ddim_net.unet_forward(sd, x, t)
    sig = lambda f: ddim_net.sigma_forward(ssd, f)
x = torch.randn(B, 3, R, R, device=dev)
t = torch.full((B,), 500.0, device=dev)


def step():
    with torch.no_grad():
        r = sig(enc(x, t))
        return fwd(x * (1 + r).reshape(-1, 1, 1, 1), t)


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
print("torch eager (cuDNN, allow_tf32=%s) %s B=%d: %.3f ms per NLC timestep, peak memory %.1f GB" % (
    torch.backends.cudnn.allow_tf32, name, B, ms, torch.cuda.max_memory_allocated() / 1e9))
