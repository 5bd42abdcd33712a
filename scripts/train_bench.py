"""The sigma-model's training forward + backward + optimizer step at the c2 shape (dim 4, 512 channels, batch 128): the native
pass (training.NativeSigmaModel: nlc_sgemm + csrc/sigma_train.cu + fused AdamW/EMA) against PyTorch autograd through the same
network (the oracle's functional forward on the GPU with TF32 allowed - what the reference's own training loop runs - and
torch.optim.AdamW + the EMA loop).  Baseline leg only: test infrastructure.
    python scripts/train_bench.py [batch]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from nlc_b200 import training as T
from oracle import adm_net, ddim_net, weights

dev = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
for name in ("c2", "c1", "adm256"):  # adm256: the c4 / c5 sigma-model (dim 8, 1024 channels, 16 heads; ADM family)
    if name.startswith("adm"):
        acfg = dict(weights.ADM_CONFIGS[name])
        cfg = acfg.pop("sigma")
        ssd = weights.adm_sigma_state_dict(**cfg, seed=4)
        extra = dict(family="adm", num_heads=acfg["num_heads"], num_head_channels=acfg["num_head_channels"],
                     use_new_attention_order=acfg["use_new_attention_order"])
        fwd = lambda sd_, f_: adm_net.sigma_forward(sd_, f_, acfg, training=True)
        Bn = min(B, 64)
    else:
        cfg = weights.CONFIGS[name]["sigma"]
        ssd = weights.ddim_sigma_state_dict(**cfg, seed=4)
        extra = {}
        fwd = lambda sd_, f_: ddim_net.sigma_forward(sd_, f_, training=True)
        Bn = B
    g = torch.Generator().manual_seed(1)
    feat = torch.randn(Bn, cfg["dim"], cfg["dim"], cfg["channels"], generator=g).to(dev)  # NHWC, as the engine hands it over
    target = (1 + 0.3 * torch.randn(Bn, generator=g)).to(dev)
    m = T.NativeSigmaModel(**cfg, dropout=0.0, loss="l2", device=dev, **extra).load_state_dict(ssd)

    def native():
        m.loss_and_grad(feat, target, nhwc=True)
        m.step(1e-4, weight_decay=0.01)

    names = [k for k in ssd if not k.endswith(("running_mean", "running_var", "num_batches_tracked"))]
    params = {n: torch.nn.Parameter(ssd[n].clone().to(dev)) for n in names}
    sd = {k: v.to(dev) for k, v in ssd.items()}
    sd.update(params)
    opt = torch.optim.AdamW(list(params.values()), lr=1e-4, weight_decay=0.01)
    ema = [p.detach().clone() for p in params.values()]
    feat_nchw = feat.permute(0, 3, 1, 2).contiguous()
    torch.backends.cudnn.allow_tf32 = True
    torch.backends.cudnn.benchmark = True

    def autograd():
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.mse_loss(fwd(sd, feat_nchw).reshape(-1) + 1, target)
        loss.backward()
        opt.step()
        for e, p in zip(ema, params.values()):
            e.mul_(0.999).add_(p.detach(), alpha=0.001)

    def timed(fn, n=10):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n

    tn, ta = timed(native), timed(autograd)
    print("%s sigma-model (dim %d, %d channels), batch %d: native fwd+bwd+AdamW/EMA %.2f ms, torch autograd (cuDNN, TF32 allowed) "
          "+ torch.optim.AdamW + EMA loop %.2f ms" % (name, cfg["dim"], cfg["channels"], Bn, tn, ta), flush=True)
