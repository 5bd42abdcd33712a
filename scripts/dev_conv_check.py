"""Development check of nlc_conv_tc on a B200: parity against torch fp32 conv on operand-rounded inputs,
plus a quick throughput number.  Run through gpurun; not part of the test-suite."""
import sys
import os
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from nlc_b200 import ops
from nlc_b200._lib import NLC_BF16, NLC_F32

torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")


def rnd(x, dt):
    if dt == NLC_BF16:
        return x.to(torch.bfloat16).float()
    return ops.round_tf32_(x.clone())


def case(name, B, H, W, Cin, Cout, dt, stride=1, pad=(1, 1, 1, 1), k=3, extras=True, time_it=False):
    g = torch.Generator(device="cpu").manual_seed(1)
    x = torch.randn(B, Cin, H, W, generator=g).to(dev)
    w = (torch.randn(Cout, Cin, k, k, generator=g) / (Cin * k * k) ** 0.5).to(dev)
    b = torch.randn(Cout, generator=g).to(dev)
    xr, wr = rnd(x, dt), rnd(w, dt)
    xp = F.pad(xr, pad)  # (left, right, top, bottom)
    ref = F.conv2d(xp, wr, b, stride=stride)
    Ho, Wo = ref.shape[2], ref.shape[3]
    tdt = ops.OP_DTYPES[dt]
    xa = ops.Act(xr.permute(0, 2, 3, 1).contiguous().to(tdt))
    wk = ops.pack_conv_weight(wr, dt)
    segs = [(0, kh - pad[2], kw - pad[0], 0, Cin) for kh in range(k) for kw in range(k)]
    rowvec = resid = None
    scale = 1.0
    if extras:
        rowvec = torch.randn(B, Cout, generator=g).to(dev)
        resid_t = torch.randn(B, Ho, Wo, Cout, generator=g).to(dev)
        resid = ops.Act(resid_t)
        scale = 0.70710678
        ref = (ref + rowvec[:, :, None, None] + resid_t.permute(0, 3, 1, 2)) * scale
    out32 = ops.Act(torch.full((B, Ho, Wo, Cout), float("nan"), device=dev))
    outop = ops.Act(torch.zeros(B, Ho, Wo, Cout, device=dev, dtype=tdt))
    ops.conv_tc([xa], segs, wk, Cout, B, Ho, Wo, dt, stride=stride, bias=b, rowvec=rowvec, resid=resid,
                out_scale=scale, out_f32=out32, out_op=outop)
    torch.cuda.synchronize()
    got = out32.t.permute(0, 3, 1, 2)
    err = (got - ref).abs().max().item()
    rel = err / ref.abs().max().item()
    err_op = (outop.t.float().permute(0, 3, 1, 2) - ref).abs().max().item() / ref.abs().max().item()
    msg = "%-34s max|err| %.3e rel %.3e  op-copy rel %.3e" % (name, err, rel, err_op)
    if time_it:
        for _ in range(3):
            ops.conv_tc([xa], segs, wk, Cout, B, Ho, Wo, dt, stride=stride, bias=b, out_f32=out32)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        n = 10
        for _ in range(n):
            ops.conv_tc([xa], segs, wk, Cout, B, Ho, Wo, dt, stride=stride, bias=b, out_f32=out32)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        fl = 2.0 * B * Ho * Wo * Cout * Cin * k * k
        msg += "  %.3f ms  %.1f TFLOP/s" % (ms, fl / ms / 1e9)
    print(msg, flush=True)
    return rel


def fused_shortcut(dt):
    """3x3 conv over source 0 plus a 1x1 shortcut over source 1 in the same accumulator."""
    B, H, W, C0, C1, Cout = 4, 16, 16, 128, 256, 128
    g = torch.Generator(device="cpu").manual_seed(2)
    x0 = rnd(torch.randn(B, C0, H, W, generator=g).to(dev), dt)
    x1 = rnd(torch.randn(B, C1, H, W, generator=g).to(dev), dt)
    w0 = rnd((torch.randn(Cout, C0, 3, 3, generator=g) / (C0 * 9) ** 0.5).to(dev), dt)
    w1 = rnd((torch.randn(Cout, C1, 1, 1, generator=g) / C1 ** 0.5).to(dev), dt)
    ref = F.conv2d(x0, w0, padding=1) + F.conv2d(x1, w1)
    tdt = ops.OP_DTYPES[dt]
    # source 1 lives in the right half of a wider buffer (ld != C)
    buf = torch.zeros(B, H, W, C1 + 64, device=dev, dtype=tdt)
    buf[..., 64:] = x1.permute(0, 2, 3, 1).to(tdt)
    a0 = ops.Act(x0.permute(0, 2, 3, 1).contiguous().to(tdt))
    a1 = ops.Act(buf, 64, C1)
    wk = ops.pack_conv_weight(w0, dt, extra=w1)
    segs = ops.taps3x3(0, 0, C0) + [(1, 0, 0, 0, C1)]
    out = ops.Act(torch.zeros(B, H, W, Cout, device=dev))
    ops.conv_tc([a0, a1], segs, wk, Cout, B, H, W, dt, out_f32=out)
    torch.cuda.synchronize()
    rel = ((out.t.permute(0, 3, 1, 2) - ref).abs().max() / ref.abs().max()).item()
    print("%-34s rel %.3e" % ("fused 3x3 + 1x1 shortcut dt=%d" % dt, rel), flush=True)
    return rel


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    worst = 0.0
    for dt in (NLC_BF16, NLC_F32):
        tag = "bf16" if dt == NLC_BF16 else "tf32"
        worst = max(worst, case("3x3 s1 16x16 128->128 B4 " + tag, 4, 16, 16, 128, 128, dt))
        worst = max(worst, case("3x3 s1 4x4 256->512 B5 " + tag, 5, 4, 4, 256, 512, dt))
        worst = max(worst, case("1x1 32x32 256->768 B3 " + tag, 3, 32, 32, 256, 768, dt, pad=(0, 0, 0, 0), k=1))
        worst = max(worst, case("3x3 s2 pad(0,1,0,1) 32->16 " + tag, 4, 32, 32, 128, 128, dt, stride=2,
                                pad=(0, 1, 0, 1)))
        worst = max(worst, case("3x3 s2 pad1 8->4 " + tag, 6, 8, 8, 256, 256, dt, stride=2))
        worst = max(worst, case("3x3 s1 256x256 64->64 B1 " + tag, 1, 256, 256, 64, 64, dt))
        worst = max(worst, fused_shortcut(dt))
        worst = max(worst, case("3x3 s1 64x64 256->256 B32 " + tag, 32, 64, 64, 256, 256, dt, extras=False,
                                time_it=True))
        worst = max(worst, case("3x3 s1 64x64 128->128 B64 " + tag, 64, 64, 64, 128, 128, dt, extras=False,
                                time_it=True))
        worst = max(worst, case("3x3 s1 16x16 512->512 B128 " + tag, 128, 16, 16, 512, 512, dt, extras=False,
                                time_it=True))
    print("WORST rel", worst)
    sys.exit(0 if worst < 2e-2 else 1)
