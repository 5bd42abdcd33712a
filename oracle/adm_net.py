"""ORACLE (test infrastructure, not product code) — CPU fp32 restatement of the reference's ADM UNet and its
sigma-model as pure functions of the reference `state_dict()`.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
Pinned live against src/unet_adm.py by tests/test_oracle_vs_reference.py and by tests/golden/nets_adm_tiny.pt.

Follows src/nn_util.py:17-19,93-121 (GroupNorm32 eps 1e-5, cos||sin timestep embedding), src/unet_adm.py:81-140
(Upsample / Downsample), :143-256 (ResBlock incl. scale-shift norm and resblock up/down), :259-305 (AttentionBlock),
:328-389 (QKVAttentionLegacy / QKVAttention), :636-693 (forward / encode), :734-796, 1029-1083 (PureResNetBlock,
SigmaModel).
"""
import math

import torch
import torch.nn.functional as F


def timestep_embedding(t, dim, max_period=10000):
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def _gn(sd, p, x):
    return F.group_norm(x.float(), 32, sd[p + ".weight"], sd[p + ".bias"], eps=1e-5)


def _conv(sd, p, x, stride=1, padding=1):
    return F.conv2d(x, sd[p + ".weight"], sd[p + ".bias"], stride=stride, padding=padding)


def res_block(sd, p, x, emb, scale_shift, updown=None):
    """src/unet_adm.py:236-256.  updown: None | 'up' | 'down' (resblock_updown: pool/upsample both h and x)."""
    h = F.silu(_gn(sd, p + "in_layers.0", x))
    if updown == "down":
        h, x = F.avg_pool2d(h, 2), F.avg_pool2d(x, 2)
    elif updown == "up":
        h, x = F.interpolate(h, scale_factor=2, mode="nearest"), F.interpolate(x, scale_factor=2, mode="nearest")
    h = _conv(sd, p + "in_layers.2", h)
    if emb is not None:
        e = F.linear(F.silu(emb), sd[p + "emb_layers.1.weight"], sd[p + "emb_layers.1.bias"])[:, :, None, None]
        if scale_shift:
            scale, shift = torch.chunk(e, 2, dim=1)
            h = F.silu(_gn(sd, p + "out_layers.0", h) * (1 + scale) + shift)
        else:
            h = F.silu(_gn(sd, p + "out_layers.0", h + e))
    else:
        h = F.silu(_gn(sd, p + "out_layers.0", h))
    h = _conv(sd, p + "out_layers.3", h)
    if p + "skip_connection.weight" in sd:
        w = sd[p + "skip_connection.weight"]
        x = F.conv2d(x, w, sd[p + "skip_connection.bias"], padding=w.shape[-1] // 2)
    return x + h


def attention_block(sd, p, x, head_channels, num_heads, new_order):
    """src/unet_adm.py:299-305 with QKVAttentionLegacy (:340-354) or QKVAttention (:373-389)."""
    b, c, hh, ww = x.shape
    T = hh * ww
    heads = c // head_channels if head_channels != -1 else num_heads
    xf = x.reshape(b, c, T)
    qkv = F.conv1d(_gn(sd, p + "norm", xf), sd[p + "qkv.weight"], sd[p + "qkv.bias"])
    ch = c // heads
    scale = 1 / math.sqrt(math.sqrt(ch))
    if new_order:
        q, k, v = qkv.chunk(3, dim=1)
        q, k, v = (z.reshape(b * heads, ch, T) for z in (q, k, v))
    else:
        q, k, v = qkv.reshape(b * heads, ch * 3, T).split(ch, dim=1)
    w = torch.einsum("bct,bcs->bts", q * scale, k * scale)
    w = torch.softmax(w.float(), dim=-1)
    a = torch.einsum("bts,bcs->bct", w, v).reshape(b, -1, T)
    h = F.conv1d(a, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])
    return (xf + h).reshape(b, c, hh, ww)


def _is_res(sd, p):
    return p + "in_layers.0.weight" in sd


def _is_attn(sd, p):
    return p + "qkv.weight" in sd


def _run_block(sd, prefix, h, emb, cfg, is_output):
    """One TimestepEmbedSequential: children are ResBlocks, AttentionBlocks, plain convs or (conv) resamplers."""
    j = 0
    while any(k.startswith("%s%d." % (prefix, j)) for k in sd):
        p = "%s%d." % (prefix, j)
        if _is_res(sd, p):
            updown = None
            if cfg["resblock_updown"] and j > 0 and is_output:
                updown = "up"
            elif cfg["resblock_updown"] and not is_output and p in cfg["_down_blocks"]:
                updown = "down"
            h = res_block(sd, p, h, emb, cfg["use_scale_shift_norm"], updown)
        elif _is_attn(sd, p):
            h = attention_block(sd, p, h, cfg["num_head_channels"], cfg["num_heads"], cfg["use_new_attention_order"])
        elif p + "op.weight" in sd:  # Downsample(use_conv=True)
            h = _conv(sd, p + "op", h, stride=2)
        elif p + "conv.weight" in sd:  # Upsample(use_conv=True)
            h = _conv(sd, p + "conv", F.interpolate(h, scale_factor=2, mode="nearest"))
        elif p + "weight" in sd:  # the input conv
            h = _conv(sd, p[:-1], h)
        j += 1
    return h


def _down_block_prefixes(cfg):
    """input_blocks indices that are ResBlock(down=True): one after each level's res blocks except the last."""
    out, idx = set(), 1
    L = len(cfg["channel_mult"])
    for level in range(L):
        idx += cfg["num_res_blocks"]
        if level != L - 1:
            out.add("input_blocks.%d.0." % idx)
            idx += 1
    return out


def _prepare(cfg):
    cfg = dict(cfg)
    cfg["_down_blocks"] = _down_block_prefixes(cfg)
    return cfg


def _embed(sd, t, cfg):
    e = timestep_embedding(t, cfg["model_channels"])
    e = F.linear(e, sd["time_embed.0.weight"], sd["time_embed.0.bias"])
    return F.linear(F.silu(e), sd["time_embed.2.weight"], sd["time_embed.2.bias"])


def _n_blocks(sd, name):
    n = 0
    while any(k.startswith("%s.%d." % (name, n)) for k in sd):
        n += 1
    return n


def unet_encode(sd, x, t, cfg):
    """src/unet_adm.py:668-693 with feat_layer = 1 (input blocks + middle block)."""
    cfg = _prepare(cfg)
    emb = _embed(sd, t, cfg)
    h = x
    for i in range(_n_blocks(sd, "input_blocks")):
        h = _run_block(sd, "input_blocks.%d." % i, h, emb, cfg, False)
    return _run_block(sd, "middle_block.", h, emb, cfg, False)


def unet_forward(sd, x, t, cfg, return_feat=False):
    """src/unet_adm.py:636-666."""
    cfg = _prepare(cfg)
    emb = _embed(sd, t, cfg)
    hs, h = [], x
    for i in range(_n_blocks(sd, "input_blocks")):
        h = _run_block(sd, "input_blocks.%d." % i, h, emb, cfg, False)
        hs.append(h)
    h = _run_block(sd, "middle_block.", h, emb, cfg, False)
    feat = h
    for i in range(_n_blocks(sd, "output_blocks")):
        h = _run_block(sd, "output_blocks.%d." % i, torch.cat([h, hs.pop()], dim=1), emb, cfg, True)
    out = _conv(sd, "out.2", F.silu(_gn(sd, "out.0", h)))
    return (out, feat) if return_feat else out


def sigma_forward(sd, feat, cfg, training=False):
    """src/unet_adm.py:1029-1083.  training=True: BatchNorm1d with batch statistics (the module in train mode, dropout 0) -
    the forward the native backward pass of nlc_b200.training.NativeSigmaModel(family="adm") is checked against."""
    h = feat
    max_idx = max(int(k.split(".")[1]) for k in sd if k.startswith("down_layer."))
    for idx in range(max_idx + 1):
        p = "down_layer.%d." % idx
        if _is_res(sd, p):
            if h.shape[-1] % 2 != 0:
                h = F.pad(h, (0, 1, 0, 1))
            h = res_block(sd, p, h, None, False)
        elif _is_attn(sd, p):
            h = attention_block(sd, p, h, cfg["num_head_channels"], cfg["num_heads"], cfg["use_new_attention_order"])
        elif p + "op.weight" in sd:
            h = _conv(sd, p + "op", h, stride=2)
    h = F.linear(h.flatten(1), sd["fc_layer.1.weight"], sd["fc_layer.1.bias"])
    if training:
        h = F.batch_norm(h, None, None, sd["fc_layer.2.weight"], sd["fc_layer.2.bias"], training=True, eps=1e-5)
    else:
        h = F.batch_norm(h, sd["fc_layer.2.running_mean"], sd["fc_layer.2.running_var"], sd["fc_layer.2.weight"],
                         sd["fc_layer.2.bias"], training=False, eps=1e-5)
    return F.linear(F.gelu(h), sd["final_mlp.weight"], sd["final_mlp.bias"])[:, :, None, None]
