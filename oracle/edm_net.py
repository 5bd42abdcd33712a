"""ORACLE (test infrastructure, not product code) — CPU fp32 restatement of the reference's EDM DDPM++ network
(`SongUNet`, second definition, the one carrying `encode`) and of the EDM sigma-model, as pure functions of the
reference `state_dict()`.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
Pinned live against src/edm_networks.py by tests/test_oracle_vs_reference.py and by tests/golden/nets_edm.pt.

Follows src/edm_networks.py:32-45 (Linear), :52-98 (Conv2d incl. the [1,1] resample filter: depthwise stride-2
box filter = 2x2 average pooling, transposed = nearest-neighbour x2), :105-116 (GroupNorm, groups =
min(32, C//4)), :124-130 (AttentionOp), :148-205 (UNetBlock), :212-225 (PositionalEmbedding), :732-909 (SongUNet
forward / encode), :912-955 (PureUNetBlock), :958-1022 (Downsample, SigmaModel).
Covers the configuration the factory builds (src/script_util.py:222-270): embedding 'positional', encoder /
decoder 'standard', adaptive_scale False, resample_proj True, num_heads 1, skip_scale sqrt(0.5), eps 1e-6.

`dhariwal_forward` restates `DhariwalUNet.forward` (src/edm_networks.py:406-502, the ADM architecture inside the EDM code
base: adaptive scale/shift, 64-channel heads, skip_scale 1, eps 1e-5, weight-less resampling skips, cos||sin embedding);
it has no `encode`, so it serves the 'base' styles only.  Pinned by tests/golden/nets_dhariwal.pt and live.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

SKIP_SCALE = np.sqrt(0.5)
EPS = 1e-6


def _gn(sd, p, x, eps=EPS):
    c = x.shape[1]
    return F.group_norm(x, min(32, c // 4), sd[p + ".weight"], sd[p + ".bias"], eps=eps)


def _conv(sd, p, x, up=False, down=False):
    """Conv2d.forward, non-fused resample branch (src/edm_networks.py:85-98)."""
    w = sd.get(p + ".weight")
    b = sd.get(p + ".bias")
    f = sd.get(p + ".resample_filter")
    c = x.shape[1]
    if up:
        x = F.conv_transpose2d(x, f.mul(4).tile([c, 1, 1, 1]), groups=c, stride=2, padding=(f.shape[-1] - 1) // 2)
    if down:
        x = F.conv2d(x, f.tile([c, 1, 1, 1]), groups=c, stride=2, padding=(f.shape[-1] - 1) // 2)
    if w is not None:
        x = F.conv2d(x, w, padding=w.shape[-1] // 2)
    if b is not None:
        x = x.add_(b.reshape(1, -1, 1, 1))
    return x


def _attention(sd, p, x):
    """num_heads = 1 (src/edm_networks.py:196-204)."""
    B, C = x.shape[:2]
    q, k, v = _conv(sd, p + "qkv", _gn(sd, p + "norm2", x)).reshape(B, C, 3, -1).unbind(2)
    w = torch.einsum("ncq,nck->nqk", q.to(torch.float32), (k / np.sqrt(k.shape[1])).to(torch.float32)).softmax(dim=2)
    a = torch.einsum("nqk,nck->ncq", w, v)
    x = _conv(sd, p + "proj", a.reshape(*x.shape)).add_(x)
    return x * SKIP_SCALE


def unet_block(sd, p, x, emb, up=False, down=False):
    """UNetBlock.forward with adaptive_scale False (src/edm_networks.py:186-205)."""
    orig = x
    x = _conv(sd, p + "conv0", F.silu(_gn(sd, p + "norm0", x)), up=up, down=down)
    params = F.linear(emb, sd[p + "affine.weight"]).add_(sd[p + "affine.bias"]).unsqueeze(2).unsqueeze(3)
    x = F.silu(_gn(sd, p + "norm1", x.add_(params)))
    x = _conv(sd, p + "conv1", x)
    has_skip = (p + "skip.weight") in sd or (p + "skip.resample_filter") in sd
    x = x.add_(_conv(sd, p + "skip", orig, up=up, down=down) if has_skip else orig)
    x = x * SKIP_SCALE
    if p + "qkv.weight" in sd:
        x = _attention(sd, p, x)
    return x


def pure_block(sd, p, x):
    """PureUNetBlock.forward (src/edm_networks.py:940-955): conv0 feeds conv1 directly (norm1 is never applied)."""
    orig = x
    x = _conv(sd, p + "conv0", F.silu(_gn(sd, p + "norm0", x)))
    x = _conv(sd, p + "conv1", x)
    x = x.add_(orig)
    x = x * SKIP_SCALE
    if p + "qkv.weight" in sd:
        x = _attention(sd, p, x)
    return x


def embedding(sd, noise_labels, model_channels):
    """PositionalEmbedding(endpoint=True), sin/cos swap, two SiLU-activated Linear layers (:837-847)."""
    half = model_channels // 2
    freqs = torch.arange(start=0, end=half, dtype=torch.float32)
    freqs = freqs / (half - 1)
    freqs = (1 / 10000) ** freqs
    x = noise_labels.ger(freqs.to(noise_labels.dtype))
    emb = torch.cat([x.cos(), x.sin()], dim=1)
    emb = emb.reshape(emb.shape[0], 2, -1).flip(1).reshape(*emb.shape)
    emb = F.silu(F.linear(emb, sd["map_layer0.weight"]).add_(sd["map_layer0.bias"]))
    return F.silu(F.linear(emb, sd["map_layer1.weight"]).add_(sd["map_layer1.bias"]))


def _enc_names(cfg):
    names = []
    for level, _ in enumerate(cfg["channel_mult"]):
        res = cfg["img_resolution"] >> level
        names.append("%dx%d_conv" % (res, res) if level == 0 else "%dx%d_down" % (res, res))
        names += ["%dx%d_block%d" % (res, res, i) for i in range(cfg["num_blocks"])]
    return names


def _dec_names(cfg):
    names = []
    L = len(cfg["channel_mult"])
    for level in reversed(range(L)):
        res = cfg["img_resolution"] >> level
        if level == L - 1:
            names += ["%dx%d_in0" % (res, res), "%dx%d_in1" % (res, res)]
        else:
            names.append("%dx%d_up" % (res, res))
        names += ["%dx%d_block%d" % (res, res, i) for i in range(cfg["num_blocks"] + 1)]
    return names


def _encoder(sd, x, emb, cfg):
    skips = []
    for name in _enc_names(cfg):
        p = "enc.%s." % name
        if name.endswith("_conv"):
            x = _conv(sd, p[:-1], x)
        else:
            x = unet_block(sd, p, x, emb, down=name.endswith("_down"))
        skips.append(x)
    return x, skips


def unet_encode(sd, x, noise_labels, cfg):
    """SongUNet.encode (src/edm_networks.py:880-909): the last encoder block's output."""
    emb = embedding(sd, noise_labels, cfg["model_channels"])
    return _encoder(sd, x, emb, cfg)[0]


def unet_forward(sd, x, noise_labels, cfg, return_feat=False):
    """SongUNet.forward (:835-878)."""
    emb = embedding(sd, noise_labels, cfg["model_channels"])
    x, skips = _encoder(sd, x, emb, cfg)
    feat = x
    for name in _dec_names(cfg):
        p = "dec.%s." % name
        if x.shape[1] != sd[p + "conv0.weight"].shape[1]:
            x = torch.cat([x, skips.pop()], dim=1)
        x = unet_block(sd, p, x, emb, up=name.endswith("_up"))
    res = cfg["img_resolution"]
    tmp = _gn(sd, "dec.%dx%d_aux_norm" % (res, res), x)
    out = _conv(sd, "dec.%dx%d_aux_conv" % (res, res), F.silu(tmp))
    return (out, feat) if return_feat else out


def sigma_forward(sd, feat, training=False):
    """SigmaModel.forward (src/edm_networks.py:1014-1022).  training=True: BatchNorm1d with batch statistics (the module in
    train mode with dropout 0): the forward nlc_b200.training.NativeSigmaModel(family="edm") is differentiated against."""
    h = feat
    max_idx = max(int(k.split(".")[1]) for k in sd if k.startswith("down_layer."))
    for idx in range(max_idx + 1):
        p = "down_layer.%d." % idx
        if p + "conv0.weight" in sd:
            if h.shape[-1] % 2 != 0:
                h = F.pad(h, (0, 1, 0, 1))
            h = pure_block(sd, p, h)
        elif p + "conv.weight" in sd:
            h = F.conv2d(F.pad(h, (0, 1, 0, 1)), sd[p + "conv.weight"], sd[p + "conv.bias"], stride=2)
    h = F.linear(h.flatten(1), sd["fc_layer.1.weight"], sd["fc_layer.1.bias"])
    if training:
        h = F.batch_norm(h, None, None, sd["fc_layer.2.weight"], sd["fc_layer.2.bias"], training=True, eps=1e-5)
    else:
        h = F.batch_norm(h, sd["fc_layer.2.running_mean"], sd["fc_layer.2.running_var"], sd["fc_layer.2.weight"],
                         sd["fc_layer.2.bias"], training=False, eps=1e-5)
    return F.linear(F.silu(h), sd["final_mlp.weight"], sd["final_mlp.bias"])[:, :, None, None]


# ------------------------------------------------------------------------------------------------ DhariwalUNet
def dhariwal_block(sd, p, x, emb, up=False, down=False):
    """UNetBlock.forward with the DhariwalUNet settings (src/edm_networks.py:186-205 with adaptive_scale True, skip_scale 1,
    eps 1e-5, channels_per_head 64, resample_proj False: a block that only resamples has a weight-less skip)."""
    orig = x
    x = _conv(sd, p + "conv0", F.silu(_gn(sd, p + "norm0", x, eps=1e-5)), up=up, down=down)
    params = F.linear(emb, sd[p + "affine.weight"]).add_(sd[p + "affine.bias"]).unsqueeze(2).unsqueeze(3)
    scale, shift = params.chunk(chunks=2, dim=1)
    x = F.silu(torch.addcmul(shift, _gn(sd, p + "norm1", x, eps=1e-5), scale + 1))
    x = _conv(sd, p + "conv1", x)
    has_skip = (p + "skip.weight") in sd or (p + "skip.resample_filter") in sd
    x = x.add_(_conv(sd, p + "skip", orig, up=up, down=down) if has_skip else orig)
    x = x * 1
    if p + "qkv.weight" in sd:
        B, C = x.shape[:2]
        heads = C // 64
        q, k, v = _conv(sd, p + "qkv", _gn(sd, p + "norm2", x, eps=1e-5)).reshape(B * heads, C // heads, 3, -1).unbind(2)
        w = torch.einsum("ncq,nck->nqk", q.to(torch.float32), (k / np.sqrt(k.shape[1])).to(torch.float32)).softmax(dim=2)
        a = torch.einsum("nqk,nck->ncq", w, v)
        x = _conv(sd, p + "proj", a.reshape(*x.shape)).add_(x)
        x = x * 1
    return x


def dhariwal_embedding(sd, noise_labels, model_channels):
    """PositionalEmbedding(endpoint=False), cos || sin, two SiLU-activated Linear layers (:475-486)."""
    half = model_channels // 2
    freqs = torch.arange(start=0, end=half, dtype=torch.float32)
    freqs = freqs / half
    freqs = (1 / 10000) ** freqs
    x = noise_labels.ger(freqs.to(noise_labels.dtype))
    emb = torch.cat([x.cos(), x.sin()], dim=1)
    emb = F.silu(F.linear(emb, sd["map_layer0.weight"]).add_(sd["map_layer0.bias"]))
    emb = F.linear(emb, sd["map_layer1.weight"]).add_(sd["map_layer1.bias"])
    return F.silu(emb)


def dhariwal_forward(sd, x, noise_labels, cfg):
    """DhariwalUNet.forward (src/edm_networks.py:475-502), unconditional (label_dim = augment_dim = 0)."""
    emb = dhariwal_embedding(sd, noise_labels, cfg["model_channels"])
    skips = []
    for name in _enc_names(cfg):
        p = "enc.%s." % name
        if name.endswith("_conv"):
            x = _conv(sd, p[:-1], x)
        else:
            x = dhariwal_block(sd, p, x, emb, down=name.endswith("_down"))
        skips.append(x)
    for name in _dec_names(cfg):
        p = "dec.%s." % name
        if x.shape[1] != sd[p + "conv0.weight"].shape[1]:
            x = torch.cat([x, skips.pop()], dim=1)
        x = dhariwal_block(sd, p, x, emb, up=name.endswith("_up"))
    return _conv(sd, "out_conv", F.silu(_gn(sd, "out_norm", x, eps=1e-5)))
