"""ORACLE (test infrastructure, not product code) — CPU fp32 restatement of the reference's DDPM/DDIM UNet and
its sigma-model, written as pure functions of the reference's `state_dict()`.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
Pinned by tests/test_oracle_vs_reference.py (against the real reference imported from /root/reference, when
present) and by the golden vectors under tests/golden/ (generated from the reference itself by
tests/golden/make_golden.py).

Follows: src/unet_ddim.py:28-46 (timestep embedding), :54-55 (GroupNorm32, eps 1e-6), :58-96 (up/down-sample),
:99-156 (ResnetBlock), :159-211 (AttnBlock), :323-393 (forward / encode), :438-529 (PureResnetBlock, SigmaModel).
"""
import math

import torch
import torch.nn.functional as F


def swish(x):
    return x * torch.sigmoid(x)


def timestep_embedding(t, dim):
    # src/unet_ddim.py:28-46: sin || cos, frequencies exp(-i * log(1e4)/(half-1))
    half = dim // 2
    k = math.log(10000) / (half - 1)
    freqs = torch.exp(torch.arange(half, dtype=torch.float32) * -k).to(t.device)
    arg = t.float()[:, None] * freqs[None, :]
    out = torch.cat([torch.sin(arg), torch.cos(arg)], dim=1)
    if dim % 2 == 1:
        out = F.pad(out, (0, 1, 0, 0))
    return out


def _gn(sd, p, x):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], eps=1e-6)


def _conv(sd, p, x, stride=1, padding=1):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=padding)


def resnet_block(sd, p, x, temb=None):
    # src/unet_ddim.py:137-156 (and :473-490 without temb)
    h = _conv(sd, p + "conv1", swish(_gn(sd, p + "norm1", x)))
    if temb is not None:
        h = h + F.linear(swish(temb), sd[p + "temb_proj.weight"], sd[p + "temb_proj.bias"])[:, :, None, None]
    h = _conv(sd, p + "conv2", swish(_gn(sd, p + "norm2", h)))
    if p + "nin_shortcut.weight" in sd:
        x = _conv(sd, p + "nin_shortcut", x, padding=0)
    elif p + "conv_shortcut.weight" in sd:
        x = _conv(sd, p + "conv_shortcut", x)
    return x + h


def attn_block(sd, p, x):
    # src/unet_ddim.py:186-211: single head over C channels, logits scaled by C^-1/2
    b, c, hh, ww = x.shape
    n = _gn(sd, p + "norm", x)
    q = _conv(sd, p + "q", n, padding=0).reshape(b, c, hh * ww)
    k = _conv(sd, p + "k", n, padding=0).reshape(b, c, hh * ww)
    v = _conv(sd, p + "v", n, padding=0).reshape(b, c, hh * ww)
    w = torch.softmax(torch.bmm(q.transpose(1, 2), k) * (int(c) ** (-0.5)), dim=2)  # [b, query, key]
    o = torch.bmm(v, w.transpose(1, 2)).reshape(b, c, hh, ww)
    return x + _conv(sd, p + "proj_out", o, padding=0)


def downsample(sd, p, x):
    # src/unet_ddim.py:89-94: zero pad right/bottom, 3x3 stride 2
    return _conv(sd, p + "conv", F.pad(x, (0, 1, 0, 1)), stride=2, padding=0)


def upsample(sd, p, x):
    return _conv(sd, p + "conv", F.interpolate(x, scale_factor=2.0, mode="nearest"))


def _levels(sd, prefix):
    n = 0
    while any(k.startswith("%s.%d." % (prefix, n)) for k in sd):
        n += 1
    return n


def _count(sd, prefix):
    n = 0
    while any(k.startswith("%s.%d." % (prefix, n)) for k in sd):
        n += 1
    return n


def _temb(sd, t):
    ch = sd["temb.dense.0.weight"].shape[1]
    e = timestep_embedding(t, ch)
    e = F.linear(e, sd["temb.dense.0.weight"], sd["temb.dense.0.bias"])
    return F.linear(swish(e), sd["temb.dense.1.weight"], sd["temb.dense.1.bias"])


def _encoder(sd, x, t):
    temb = _temb(sd, t)
    hs = [_conv(sd, "conv_in", x)]
    n_levels = _levels(sd, "down")
    for lv in range(n_levels):
        n_blocks = _count(sd, "down.%d.block" % lv)
        has_attn = _count(sd, "down.%d.attn" % lv) > 0
        for ib in range(n_blocks):
            h = resnet_block(sd, "down.%d.block.%d." % (lv, ib), hs[-1], temb)
            if has_attn:
                h = attn_block(sd, "down.%d.attn.%d." % (lv, ib), h)
            hs.append(h)
        if "down.%d.downsample.conv.weight" % lv in sd:
            hs.append(downsample(sd, "down.%d.downsample." % lv, hs[-1]))
    h = resnet_block(sd, "mid.block_1.", hs[-1], temb)
    h = attn_block(sd, "mid.attn_1.", h)
    return h, hs, temb


def unet_encode(sd, x, t, feat_layer=0):
    """src/unet_ddim.py:365-393; feat_layer 1 (src/unet_simple.py:373-375): the feature is mid.block_2's output."""
    h, _, temb = _encoder(sd, x, t)
    return h if feat_layer == 0 else resnet_block(sd, "mid.block_2.", h, temb)


def unet_forward(sd, x, t, return_feat=False, feat_layer=0):
    """src/unet_ddim.py:323-363 (and forward_and_encode :395-436 when return_feat; src/unet_simple.py:399-407 for
    feat_layer 1)."""
    feat, hs, temb = _encoder(sd, x, t)
    h = resnet_block(sd, "mid.block_2.", feat, temb)
    if feat_layer != 0:
        feat = h
    for lv in reversed(range(_levels(sd, "up"))):
        n_blocks = _count(sd, "up.%d.block" % lv)
        has_attn = _count(sd, "up.%d.attn" % lv) > 0
        for ib in range(n_blocks):
            h = resnet_block(sd, "up.%d.block.%d." % (lv, ib), torch.cat([h, hs.pop()], dim=1), temb)
            if has_attn:
                h = attn_block(sd, "up.%d.attn.%d." % (lv, ib), h)
        if "up.%d.upsample.conv.weight" % lv in sd:
            h = upsample(sd, "up.%d.upsample." % lv, h)
    out = _conv(sd, "conv_out", swish(_gn(sd, "norm_out", h)))
    return (out, feat) if return_feat else out


def sigma_forward(sd, feat, training=False):
    """SigmaModel, src/unet_ddim.py:493-529: per block [pad if odd] -> PureResnetBlock -> (AttnBlock in block 0)
    -> Downsample; Flatten -> Linear -> BatchNorm1d (eval) -> GELU -> Linear -> [B,1,1,1].
    training=True: BatchNorm1d normalises with the batch statistics (`sigma_model.train()`, src/experiments.py:643;
    dropout is taken as 0, the running statistics are not updated here) — the forward the training step differentiates."""
    h = feat
    idx = 0
    n_layers = _count(sd, "down_layer") if any(k.startswith("down_layer.") for k in sd) else 0
    # module indices with parameters are a subset (Identity/ConstantPad2d hold none): walk all indices
    max_idx = max(int(k.split(".")[1]) for k in sd if k.startswith("down_layer."))
    while idx <= max_idx:
        p = "down_layer.%d." % idx
        if p + "norm1.weight" in sd:
            if h.shape[-1] % 2 != 0:  # the ConstantPad2d slot precedes this block
                h = F.pad(h, (0, 1, 0, 1))
            h = resnet_block(sd, p, h, None)
        elif p + "q.weight" in sd:
            h = attn_block(sd, p, h)
        elif p + "conv.weight" in sd:
            h = downsample(sd, p, h)
        idx += 1
    h = h.flatten(1)
    h = F.linear(h, sd["fc_layer.1.weight"], sd["fc_layer.1.bias"])
    if training:
        h = F.batch_norm(h, None, None, sd["fc_layer.2.weight"], sd["fc_layer.2.bias"], training=True, eps=1e-5)
    else:
        h = F.batch_norm(h, sd["fc_layer.2.running_mean"], sd["fc_layer.2.running_var"], sd["fc_layer.2.weight"],
                         sd["fc_layer.2.bias"], training=False, eps=1e-5)
    h = F.gelu(h)
    out = F.linear(h, sd["final_mlp.weight"], sd["final_mlp.bias"])
    return out[:, :, None, None]
