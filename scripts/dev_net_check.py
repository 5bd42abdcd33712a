"""Development check on a B200: per-op parity vs torch, then the DDIM UNet + sigma-model vs the CPU oracle."""
import os
import sys
import time
import traceback

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nlc_b200 import ops
from nlc_b200._lib import NLC_BF16, NLC_F32
from nlc_b200.ops import Act
from oracle import ddim_net, weights

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)


def rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def check(name, fn):
    try:
        r = fn()
        print("%-44s %s" % (name, r), flush=True)
    except Exception:
        print("%-44s FAILED" % name)
        traceback.print_exc()


def t_groupnorm(dt, B=3, H=8, W=8, C=256, silu=True, ss=False):
    x = torch.randn(B, C, H, W, generator=g).to(dev) * 2 + 0.5
    gam, bet = torch.randn(C, generator=g).to(dev), torch.randn(C, generator=g).to(dev)
    ref = F.group_norm(x, 32, gam, bet, eps=1e-6)
    sc = sh = None
    if ss:
        sc, sh = torch.randn(B, C, generator=g).to(dev), torch.randn(B, C, generator=g).to(dev)
        ref = ref * (1 + sc[:, :, None, None]) + sh[:, :, None, None]
    if silu:
        ref = F.silu(ref)
    xa = Act(x.permute(0, 2, 3, 1).contiguous())
    y = Act(torch.zeros(B, H, W, C, device=dev, dtype=ops.OP_DTYPES[dt]))
    ws = torch.zeros(ops.groupnorm_ws(B, H * W, C, 32), device=dev)
    ops.groupnorm(xa, 32, 1e-6, gam, bet, y, dt, ws, silu=silu, scale=sc, shift=sh)
    return "rel %.2e" % rel(y.t.float().permute(0, 3, 1, 2), ref)


def t_attention(dt, B, T, heads, dh, legacy=False):
    C = heads * dh
    tdt = ops.OP_DTYPES[dt]
    qkv = torch.randn(B, T, 3 * C, generator=g).to(dev).to(tdt)
    f = qkv.float()
    if legacy:  # per head [q,k,v]
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
    qa = Act(qkv.view(B, side, T // side, 3 * C))
    out = Act(torch.zeros(B, side, T // side, C, device=dev, dtype=tdt))
    ws = torch.zeros(max(ops.attention_ws(dt, B, T, heads, dh), 16), device=dev, dtype=torch.uint8)
    ops.attention(qa, dt, offs[0], offs[1], offs[2], offs[3], heads, dh, scale, out, ws)
    return "rel %.2e" % rel(out.t.float().view(B, T, C), ref)


def t_linear():
    x = torch.randn(37, 300, generator=g).to(dev)
    W = torch.randn(130, 300, generator=g).to(dev) / 17
    b = torch.randn(130, generator=g).to(dev)
    y = torch.zeros(37, 130, device=dev)
    ops.linear(x, W, b, y, act_in=1, act_out=2)
    ref = F.gelu(F.linear(F.silu(x), W, b))
    return "rel %.2e" % rel(y, ref)


def t_resample(dt, mode):
    x = torch.randn(2, 64, 8, 8, generator=g).to(dev)
    ref = x if mode == 0 else (F.interpolate(x, scale_factor=2.0, mode="nearest") if mode == 1 else F.avg_pool2d(x, 2))
    xa = Act(x.permute(0, 2, 3, 1).contiguous())
    B, C, Ho, Wo = ref.shape
    yf = Act(torch.zeros(B, Ho, Wo, C, device=dev))
    yo = Act(torch.zeros(B, Ho, Wo, C, device=dev, dtype=ops.OP_DTYPES[dt]))
    ops.resample(xa, mode, yf, yo, dt)
    return "f32 rel %.2e op rel %.2e" % (rel(yf.t.permute(0, 3, 1, 2), ref), rel(yo.t.float().permute(0, 3, 1, 2), ref))


def t_conv_io(dt):
    B, R, C = 3, 16, 128
    x = torch.randn(B, 3, R, R, generator=g).to(dev)
    sc = torch.rand(B, generator=g).to(dev) + 0.5
    w = torch.randn(C, 3, 3, 3, generator=g).to(dev) / 5
    b = torch.randn(C, generator=g).to(dev)
    ref = F.conv2d(x * sc[:, None, None, None], w, b, padding=1)
    yf = Act(torch.zeros(B, R, R, C, device=dev))
    yo = Act(torch.zeros(B, R, R, C, device=dev, dtype=ops.OP_DTYPES[dt]))
    ops.conv_in_nchw(x, sc, w, b, yf, yo, dt)
    r1 = rel(yf.t.permute(0, 3, 1, 2), ref)
    w2 = torch.randn(3, C, 3, 3, generator=g).to(dev) / 30
    b2 = torch.randn(3, generator=g).to(dev)
    ref2 = F.conv2d(yo.t.float().permute(0, 3, 1, 2), w2, b2, padding=1)
    out = torch.zeros(B, 3, R, R, device=dev)
    ops.conv_out_nchw(yo, dt, w2, b2, out)
    return "conv_in rel %.2e conv_out rel %.2e" % (r1, rel(out, ref2))


def t_sampler():
    B, d = 5, 3 * 16 * 16
    x = torch.randn(B, 3, 16, 16, generator=g).to(dev) * 3
    nrm = torch.zeros(B, device=dev)
    ops.row_norm(x, nrm)
    r1 = rel(nrm, torch.linalg.vector_norm(x.view(B, -1), dim=1))
    y = x.clone()
    ops.normalize_rows_(y)
    ref = (d ** 0.5) * x / torch.linalg.vector_norm(x.view(B, -1), dim=1).clamp_min(1e-12).view(B, 1, 1, 1)
    return "norm rel %.2e normalize rel %.2e" % (r1, rel(y, ref))


def t_net(name, precision, B=2):
    from nlc_b200.unet_ddim import SigmaModel, UNetModel
    cfg = weights.CONFIGS[name]
    sd = weights.ddim_unet_state_dict(**cfg["unet"], seed=3)
    ssd = weights.ddim_sigma_state_dict(**cfg["sigma"], seed=4)
    R = cfg["unet"]["image_size"]
    x = torch.randn(B, 3, R, R, generator=g)
    t = torch.tensor([500.0, 37.0, 999.0, 0.0][:B])
    with torch.no_grad():
        t0 = time.time()
        ref, feat = ddim_net.unet_forward(sd, x, t, return_feat=True)
        r_ref = ddim_net.sigma_forward(ssd, feat)
        t_cpu = time.time() - t0
    m = UNetModel(**cfg["unet"], precision=precision, device=dev).load_state_dict(sd)
    s = SigmaModel(**cfg["sigma"], precision=precision, device=dev).load_state_dict(ssd)
    out = m(x.to(dev), t.to(dev))
    f = m.encode(x.to(dev), t.to(dev))
    r = s(f)
    r_tf = s(feat.to(dev))  # teacher-forced sigma head
    torch.cuda.synchronize()
    return "fwd rel %.2e feat rel %.2e r abs %.2e (teacher-forced %.2e; r=%s ref=%s) cpu %.2fs" % (
        rel(out.cpu(), ref), rel(f.cpu(), feat), (r.cpu() - r_ref).abs().max().item(),
        (r_tf.cpu() - r_ref).abs().max().item(), [round(v, 4) for v in r.flatten().tolist()],
        [round(v, 4) for v in r_ref.flatten().tolist()], t_cpu)


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    for dt, tag in ((NLC_BF16, "bf16"), (NLC_F32, "tf32")):
        check("groupnorm silu " + tag, lambda: t_groupnorm(dt))
        check("groupnorm C=384 HW=4096 " + tag, lambda: t_groupnorm(dt, B=2, H=64, W=64, C=384))
        check("groupnorm scale-shift no-silu " + tag, lambda: t_groupnorm(dt, silu=False, ss=True))
        check("groupnorm 2x2 C=512 " + tag, lambda: t_groupnorm(dt, B=5, H=2, W=2, C=512))
        check("attention T=16 1x512 " + tag, lambda: t_attention(dt, 3, 16, 1, 512))
        check("attention T=64 4x64 legacy " + tag, lambda: t_attention(dt, 2, 64, 4, 64, legacy=True))
        check("attention T=256 1x256 " + tag, lambda: t_attention(dt, 3, 256, 1, 256))
        check("attention T=1024 4x64 " + tag, lambda: t_attention(dt, 2, 1024, 4, 64))
        check("attention T=256 4x64 legacy " + tag, lambda: t_attention(dt, 2, 256, 4, 64, legacy=True))
        check("resample copy " + tag, lambda: t_resample(dt, 0))
        check("resample up2 " + tag, lambda: t_resample(dt, 1))
        check("resample avgpool " + tag, lambda: t_resample(dt, 2))
        check("conv_in / conv_out " + tag, lambda: t_conv_io(dt))
    check("linear", t_linear)
    check("row_norm / normalize", t_sampler)
    for prec in ("tf32", "bf16"):
        check("net tiny " + prec, lambda: t_net("tiny", prec))
        check("net c1 " + prec, lambda: t_net("c1", prec))
        check("net c2 " + prec, lambda: t_net("c2", prec))
