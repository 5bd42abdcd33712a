"""Host-side mirror of the reference's EDM DDPM++ network (`SongUNet`, the definition that carries `encode`,
src/edm_networks.py:732-909) and of the EDM sigma-model (:979-1022), executing on libnlc_b200 kernels.

Constructors take the reference's arguments and consume the reference `state_dict()` unchanged (including the
`resample_filter` buffers, which are checked to be the [1,1] box filter and otherwise unused).  Calls follow the
reference: `model(x, noise_labels, class_labels=None)` -> `[B, out, R, R]`, `model.encode(...)` -> the last
encoder block's output `[B, C, h, w]`, `sigma_model(feat)` -> `[B,1,1,1]`.  `forward_scaled` / `encode_scaled`
fold the EDM input preconditioning c_in = 1/sqrt(sigma_data^2 + sigma^2) into the input convolution.

Covered configuration = what `create_edm_sigma_eps_model` builds (src/script_util.py:222-270): positional
embedding, standard encoder/decoder, resample_filter [1,1] (depthwise stride-2 box filter == 2x2 average pooling;
its transpose == nearest-neighbour x2, src/edm_networks.py:85-93), adaptive_scale False, num_heads 1,
skip_scale sqrt(0.5), GroupNorm eps 1e-6 with min(32, C//4) groups.
"""
import math

import numpy as np
import torch

from . import ops
from .engine import (Engine, Feat, PlanCtx, h_feat, emit_attention, emit_conv1x1, emit_conv3x3, emit_conv_in, emit_conv_out,
                     emit_groupnorm, emit_upsample_conv3x3, run, upsample_conv_eligible)
from .ops import Act

GN_EPS = 1e-6
SKIP_SCALE = float(np.sqrt(0.5))


def _groups(c):
    return min(32, c // 4)


def _g(sd, key):
    if key not in sd:
        raise KeyError("state_dict is missing %r" % key)
    return sd[key]


def _check_filter(sd, key):
    f = sd.get(key)
    if f is not None and not torch.equal(f.float().reshape(-1), torch.full((4,), 0.25)):
        raise NotImplementedError("%s: only resample_filter [1,1] (DDPM++) is on the path" % key)


class _BlockW:
    """UNetBlock / PureUNetBlock weights (src/edm_networks.py:148-205, 912-955)."""

    def __init__(self, eng, sd, p, up=False, down=False, with_emb=True, adaptive=False, head_ch=None, skip_scale=SKIP_SCALE,
                 eps=GN_EPS):
        """`adaptive`, `head_ch`, `skip_scale`, `eps`: the UNetBlock options in which DhariwalUNet differs from the DDPM++
        network (adaptive scale/shift, 64-channel heads, skip_scale 1, eps 1e-5; src/edm_networks.py:427-428)."""
        g = lambda k: _g(sd, p + k)
        w0 = g("conv0.weight")
        self.cin, self.cout = w0.shape[1], w0.shape[0]
        self.up, self.down, self.with_emb = up, down, with_emb
        self.adaptive, self.skip_scale, self.eps = adaptive, float(skip_scale), eps
        self.heads = 1 if head_ch is None else self.cout // head_ch
        self.resid_mode = 0
        _check_filter(sd, p + "conv0.resample_filter")
        _check_filter(sd, p + "skip.resample_filter")
        self.n0w, self.n0b = eng.dev32(g("norm0.weight")), eng.dev32(g("norm0.bias"))
        self.w0, self.b0 = eng.pack3x3(w0), eng.dev32(g("conv0.bias"))
        # up blocks: conv0 runs at the low resolution as four sub-pixel phases (engine.emit_upsample_conv3x3)
        self.w0_phase = ops.upsample_phase_weights(w0.to(eng.device), eng.op_dtype) if (up and with_emb) else None
        if with_emb:  # PureUNetBlock never applies norm1 (src/edm_networks.py:940-944)
            self.n1w, self.n1b = eng.dev32(g("norm1.weight")), eng.dev32(g("norm1.bias"))
        b1 = g("conv1.bias").float()
        self.fused_skip = (p + "skip.weight") in sd
        if self.fused_skip:
            self.w1 = eng.pack3x3(g("conv1.weight"), extra=g("skip.weight"))
            b1 = b1 + g("skip.bias").float()
        else:
            # a same-width resampling block without resample_proj has a weight-less skip (Conv2d(kernel=0), :172-176): the
            # residual is x itself, nearest x2 / 2x2 averaged - read at its own resolution by the second conv's epilogue
            self.resid_mode = 1 if up else (2 if down else 0)
            self.w1 = eng.pack3x3(g("conv1.weight"))
        self.b1 = eng.dev32(b1)
        self.aff_w = g("affine.weight") if with_emb else None
        self.aff_b = g("affine.bias").float().cpu() if with_emb else None
        if with_emb and not adaptive:
            # (the conv0 bias rides on the per-sample affine row: one vector instead of two in the conv epilogue)
            self.aff_b = self.aff_b + g("conv0.bias").float().cpu()
        self.emb_off = 0
        self.attn = (p + "qkv.weight") in sd
        if self.attn:
            C = self.cout
            self.n2w, self.n2b = eng.dev32(g("norm2.weight")), eng.dev32(g("norm2.bias"))
            # qkv output channel c*3 + j is (q,k,v)[j] of channel c (reshape(..., C, 3, T).unbind(2), num_heads 1):
            # regroup the rows as [q | k | v]
            wq, bq = g("qkv.weight"), g("qkv.bias")
            perm = torch.arange(3 * C).view(C, 3).t().reshape(-1)
            self.wqkv = eng.pack3x3(wq[perm])
            self.bqkv = eng.dev32(bq[perm])
            self.wproj, self.bproj = eng.pack3x3(g("proj.weight")), eng.dev32(g("proj.bias"))


def _emit_block(pc, w, x, dest, emb=None):
    """x: Feat with fp32 (and the operand copy when the block has a non-resampled 1x1 skip)."""
    eng = pc.eng
    dt = eng.op_dtype
    B, H, W = x.B, x.H, x.W
    skip_src = x.op
    if not (w.up or w.down):
        a0 = eng.act_op("ub.a0", B, H, W, w.cin)
        emit_groupnorm(pc, x.res, w.n0w, w.n0b, _groups(w.cin), w.eps, a0, silu=True)
    else:
        # Conv2d(up/down) with the [1,1] filter (src/edm_networks.py:85-93): nearest x2 / 2x2 average of the
        # activated tensor, written directly by the GroupNorm apply pass; the skip branch resamples x itself
        mode = 1 if w.up else 2
        src32 = x.res
        up_low = w.up and w.with_emb and upsample_conv_eligible(H, W)
        if up_low:  # the activated tensor stays at the low resolution; conv0 computes conv(upsample(a0)) from it
            a0 = eng.act_op("ub.a0", B, H, W, w.cin)
            emit_groupnorm(pc, src32, w.n0w, w.n0b, _groups(w.cin), w.eps, a0, silu=True)
        H, W = (2 * H, 2 * W) if w.up else (H // 2, W // 2)
        if not up_low:
            a0 = eng.act_op("ub.a0r", B, H, W, w.cin)
            emit_groupnorm(pc, src32, w.n0w, w.n0b, _groups(w.cin), w.eps, a0, silu=True, resample=mode)
        skip_src = None
        if w.fused_skip:  # (a weight-less skip reads x at its own resolution in the second conv's epilogue instead)
            xs = eng.act_op("ub.xs", B, H, W, w.cin)
            if src32.dtype == torch.float32:
                pc.add(lambda: ops.resample(src32, mode, None, xs, dt), "resample")
            else:  # 16-bit residual stream
                pc.add(lambda: ops.resample_op(src32, mode, xs, dt), "resample")
            skip_src = xs
    rowvec = scale = shift = None
    if emb is not None and w.with_emb:
        if w.adaptive:  # scale, shift = params.chunk(2): silu(shift + norm1(x) * (scale + 1)), :189-191
            scale = emb[:, w.emb_off:w.emb_off + w.cout]
            shift = emb[:, w.emb_off + w.cout:w.emb_off + 2 * w.cout]
        else:
            rowvec = emb[:, w.emb_off:w.emb_off + w.cout]
    res_dest = dest
    if w.attn:
        res_dest = eng.stream_feat("ub.y", B, H, W, w.cout)
    if w.with_emb:
        h = eng.act_h("ub.h", B, H, W, w.cout)
        if (w.up or w.down) and up_low:
            emit_upsample_conv3x3(pc, a0, w.w0_phase, w.b0 if rowvec is None else None, w.cout, h_feat(h), rowvec=rowvec)
        else:
            emit_conv3x3(pc, a0, w.w0, w.b0 if rowvec is None else None, w.cout, h_feat(h), rowvec=rowvec)
        a1 = eng.act_op("ub.a1", B, H, W, w.cout)
        emit_groupnorm(pc, h, w.n1w, w.n1b, _groups(w.cout), w.eps, a1, silu=True, scale=scale, shift=shift)
    else:
        a1 = eng.act_op("ub.a1", B, H, W, w.cout)
        emit_conv3x3(pc, a0, w.w0, w.b0, w.cout, Feat(op=a1))
    if w.fused_skip:
        assert skip_src is not None, "block with a 1x1 skip needs the operand copy of its input"
        emit_conv3x3(pc, a1, w.w1, w.b1, w.cout, res_dest, extra_src=skip_src, out_scale=w.skip_scale)
    else:
        emit_conv3x3(pc, a1, w.w1, w.b1, w.cout, res_dest, resid=x.res, out_scale=w.skip_scale, resid_mode=w.resid_mode)
    if w.attn:
        C = w.cout
        y = res_dest
        a2 = eng.act_op("ub.a2", B, H, W, C)
        emit_groupnorm(pc, y.res, w.n2w, w.n2b, _groups(C), w.eps, a2, silu=False)
        qkv = eng.act_op("ub.qkv", B, H, W, 3 * C)
        emit_conv1x1(pc, a2, w.wqkv, w.bqkv, 3 * C, Feat(op=qkv))
        o = eng.act_op("ub.o", B, H, W, C)
        dh = C // w.heads  # rows regrouped as [q | k | v] x [head][channel]; weights softmax(q . k / sqrt(dh)), :124-127
        emit_attention(pc, qkv, 0, C, 2 * C, dh if w.heads > 1 else 0, w.heads, dh, float(1.0 / math.sqrt(dh)), o)
        emit_conv1x1(pc, o, w.wproj, w.bproj, C, dest, resid=y.res, out_scale=w.skip_scale)


class SongUNet:
    """Drop-in for src/edm_networks.py:732 `SongUNet` (inference, unconditional DDPM++ configuration)."""
    # what DhariwalUNet (below) changes
    _BLOCK_KW = {}
    _EMB_ENDPOINT, _EMB_COS_FIRST = True, False  # PositionalEmbedding(endpoint=True) + the sin/cos swap of :838
    _OUT_NORM, _OUT_CONV = "dec.%dx%d_aux_norm", "dec.%dx%d_aux_conv"
    _OUT_EPS = GN_EPS

    def __init__(self, img_resolution, in_channels, out_channels, label_dim=0, augment_dim=0, model_channels=128,
                 channel_mult=(1, 2, 2, 2), channel_mult_emb=4, num_blocks=4, attn_resolutions=(16,), dropout=0.10,
                 label_dropout=0, embedding_type="positional", channel_mult_noise=1, encoder_type="standard",
                 decoder_type="standard", resample_filter=(1, 1), precision="bf16", device="cuda", **kwargs):
        if label_dim or augment_dim:
            raise NotImplementedError("class / augment conditioning is not on the sampling path (label_dim=0)")
        if embedding_type != "positional" or encoder_type != "standard" or decoder_type != "standard" or \
                channel_mult_noise != 1 or list(resample_filter) != [1, 1]:
            raise NotImplementedError("only the DDPM++ configuration built by create_edm_sigma_eps_model is covered")
        self.img_resolution, self.in_channels, self.out_channels = img_resolution, in_channels, out_channels
        self.model_channels, self.channel_mult = model_channels, tuple(channel_mult)
        self.emb_ch = model_channels * channel_mult_emb
        self.num_blocks, self.attn_resolutions = num_blocks, tuple(attn_resolutions)
        self.eng = Engine(device, precision)
        self._plans, self._loaded = {}, False

    def load_state_dict(self, sd, strict=True):
        eng = self.eng
        sd = {k: v.detach() for k, v in sd.items()}
        R, mc = self.img_resolution, self.model_channels
        self.m0w, self.m0b = eng.dev32(sd["map_layer0.weight"]), eng.dev32(sd["map_layer0.bias"])
        self.m1w, self.m1b = eng.dev32(sd["map_layer1.weight"]), eng.dev32(sd["map_layer1.bias"])
        half = mc // 2
        # PositionalEmbedding (src/edm_networks.py:220-224); SongUNet: endpoint=True and the sin/cos swap of :838 -> sin || cos
        freqs = torch.arange(start=0, end=half, dtype=torch.float32) / (half - (1 if self._EMB_ENDPOINT else 0))
        self.freqs = ((1 / 10000) ** freqs).to(eng.device)
        kw = self._BLOCK_KW
        p = "enc.%dx%d_conv." % (R, R)
        self.cin_w, self.cin_b = eng.dev32(sd[p + "weight"]), eng.dev32(sd[p + "bias"])
        self.cin_wp = ops.pack_conv_in_weight(self.cin_w, eng.op_dtype)
        self.enc, self.dec = [], []  # lists of (res, _BlockW, concat?)
        L = len(self.channel_mult)
        for level in range(L):
            res = R >> level
            if level > 0:
                self.enc.append((res, _BlockW(eng, sd, "enc.%dx%d_down." % (res, res), down=True, **kw)))
            for idx in range(self.num_blocks):
                self.enc.append((res, _BlockW(eng, sd, "enc.%dx%d_block%d." % (res, res, idx), **kw)))
        for level in reversed(range(L)):
            res = R >> level
            if level == L - 1:
                self.dec.append((res, _BlockW(eng, sd, "dec.%dx%d_in0." % (res, res), **kw), False))
                self.dec.append((res, _BlockW(eng, sd, "dec.%dx%d_in1." % (res, res), **kw), False))
            else:
                self.dec.append((res, _BlockW(eng, sd, "dec.%dx%d_up." % (res, res), up=True, **kw), False))
            for idx in range(self.num_blocks + 1):
                self.dec.append((res, _BlockW(eng, sd, "dec.%dx%d_block%d." % (res, res, idx), **kw), True))
        pn, pc_ = (self._OUT_NORM % (R, R) if "%" in self._OUT_NORM else self._OUT_NORM,
                   self._OUT_CONV % (R, R) if "%" in self._OUT_CONV else self._OUT_CONV)
        self.no_w, self.no_b = eng.dev32(sd[pn + ".weight"]), eng.dev32(sd[pn + ".bias"])
        self.cout_w, self.cout_b = eng.dev32(sd[pc_ + ".weight"]), eng.dev32(sd[pc_ + ".bias"])
        self.cout_packed = (ops.pack_conv_out_weight(self.cout_w, self.cout_b, eng.op_dtype)
                            if eng.chunk == 64 and self.cout_w.shape[0] <= 8 else None)
        order = [b for _, b in self.enc]
        n_enc = len(order)
        order += [b for _, b, _ in self.dec]
        off = 0
        for i, b in enumerate(order):
            if i == n_enc:
                self.emb_enc = off
            b.emb_off = off
            off += b.aff_w.shape[0]  # cout, or 2 * cout with adaptive scale/shift
        self.emb_total = off
        self.aw = eng.dev32(torch.cat([b.aff_w for b in order], dim=0))
        self.ab = eng.dev32(torch.cat([b.aff_b for b in order], dim=0))
        for b in order:
            b.aff_w = b.aff_b = None
        self._loaded, self._plans = True, {}
        return self

    @classmethod
    def from_reference(cls, m, precision="bf16", device="cuda"):
        """Build from an instance of the reference's src.edm_networks.SongUNet."""
        names = list(m.enc.keys())
        R = int(names[0].split("x")[0])
        mc = m.enc[names[0]].out_channels
        levels = sorted({int(n.split("x")[0]) for n in names}, reverse=True)
        mult, attn = [], []
        for res in levels:
            blocks = [n for n in names if n.startswith("%dx%d_block" % (res, res))]
            mult.append(m.enc[blocks[-1]].out_channels // mc)
            if m.enc[blocks[0]].num_heads:
                attn.append(res)
        nb = len([n for n in names if n.startswith("%dx%d_block" % (R, R))])
        self = cls(R, m.enc[names[0]].in_channels, m.dec["%dx%d_aux_conv" % (R, R)].out_channels, model_channels=mc,
                   channel_mult=mult, num_blocks=nb, attn_resolutions=attn,
                   channel_mult_emb=m.map_layer0.out_features // mc, precision=precision, device=device)
        return self.load_state_dict(m.state_dict())

    # ------------------------------------------------------------------ plan
    def _plan(self, B):
        if B not in self._plans:
            assert self._loaded, "load_state_dict() first"
            self._plans[B] = self.eng.plan_two_pass(lambda: self._build_plan(B))
        return self._plans[B]

    def _build_plan(self, B):
        eng, R, mc = self.eng, self.img_resolution, self.model_channels
        f32, opt, dt = torch.float32, eng.op_torch, eng.op_dtype
        P = {}
        P["x"] = eng.named("x", (B, self.in_channels, R, R), f32)
        P["t"] = eng.named("t", (B,), f32)
        P["in_scale"] = eng.named("in_scale", (B,), f32)
        P["emb_sin"] = eng.named("emb_sin", (B, mc), f32)
        P["emb_h"] = eng.named("emb_h", (B, self.emb_ch), f32)
        P["emb"] = eng.named("emb", (B, self.emb_ch), f32)
        P["aff"] = eng.named("aff", (B, self.emb_total), f32)
        P["out"] = eng.named("out", (B, self.out_channels, R, R), f32)
        # skips: the input conv, then every encoder block (src/edm_networks.py:852-862)
        skip_ch = [mc] + [b.cout for _, b in self.enc]
        skip_res = [R] + [res for res, _ in self.enc]
        n_skips = len(skip_ch)
        consumers = [b for _, b, cat in self.dec if cat]
        assert len(consumers) == n_skips
        cat = {}
        for b, k in zip(consumers, reversed(range(n_skips))):
            c_h = b.cin - skip_ch[k]
            r = skip_res[k]
            cat[k] = eng.cat_buffers(k, B, r, r, b.cin) + (c_h,)

        def skip_feat(k):
            c32, c16, c1 = cat[k]
            return eng.cat_view(c32, c16, c1, skip_ch[k])

        def head_feat(k):
            c32, c16, c1 = cat[k]
            return eng.cat_view(c32, c16, 0, c1)

        aff = P["aff"]
        enc = PlanCtx(eng, B)
        P["emb_n"], P["use_scale"] = [self.emb_enc], [False]
        x_in, in_scale = P["x"], P["in_scale"]
        enc.add(lambda: ops.timestep_embedding(P["t"], self.freqs, self._EMB_COS_FIRST, P["emb_sin"]), "embedding")
        enc.add(lambda: ops.linear(P["emb_sin"], self.m0w, self.m0b, P["emb_h"], act_out=1), "map_layer0")
        enc.add(lambda: ops.linear(P["emb_h"], self.m1w, self.m1b, P["emb"], act_out=1), "map_layer1")
        enc.add(lambda: ops.linear(P["emb"], self.aw[:P["emb_n"][0]], self.ab[:P["emb_n"][0]],
                                   aff[:, :P["emb_n"][0]]), "affine (all blocks)")
        d0 = skip_feat(0)
        emit_conv_in(enc, x_in, lambda: in_scale if P["use_scale"][0] else None, self.cin_wp, self.cin_b,
                     self.cin_w.shape[0], d0, w_f32=self.cin_w)
        cur = d0
        feat32 = Act(eng.named("feat", (B, skip_res[-1], skip_res[-1], skip_ch[-1]), f32))
        for k, (res, w) in enumerate(self.enc, start=1):
            dest = skip_feat(k)
            if k == n_skips - 1 and dest.f32 is None:
                # 16-bit residual stream at the last level: the block writes the fp32 feature tensor itself, next to the
                # operand copy in the concat buffer
                dest = Feat(f32=feat32, op=dest.op)
            _emit_block(enc, w, cur, dest, aff)
            cur = dest
        last_skip = cur
        if last_skip.f32 is not feat32:
            src32 = last_skip.f32
            enc.add(lambda: ops.resample(src32, 0, feat32, None, dt), "feat copy")
        P["feat"] = feat32.t

        dec = PlanCtx(eng, B)
        dec._gn_ws_floats, dec._attn_ws_bytes = enc._gn_ws_floats, enc._attn_ws_bytes
        k = n_skips - 1
        # decoder input x is the last encoder output; the in0/in1/up blocks chain through the concat heads
        for i, (res, w, is_cat) in enumerate(self.dec):
            if is_cat:
                c32, c16, _ = cat[k]
                x = eng.cat_view(c32, c16)
                k -= 1
            else:
                x = cur
            nxt = self.dec[i + 1] if i + 1 < len(self.dec) else None
            if nxt is None:
                dest = eng.stream_feat("dec.out", B, R, R, w.cout)
            elif nxt[2]:
                dest = head_feat(k)
            else:  # next block takes this output alone (in1 after in0, or an up block)
                dest = eng.stream_feat("dec.t%d" % (i % 2), B, res, res, w.cout, need_op=True)
            _emit_block(dec, w, x, dest, aff)
            cur = dest
        a = eng.act_op("ub.a0", B, R, R, cur.C)
        emit_groupnorm(dec, cur.res, self.no_w, self.no_b, _groups(cur.C), self._OUT_EPS, a, silu=True)
        emit_conv_out(dec, a, self.cout_w, self.cout_b, self.cout_packed, P["out"])
        m = max(enc._gn_ws_floats, dec._gn_ws_floats)
        enc._gn_ws_floats = dec._gn_ws_floats = m
        m = max(enc._attn_ws_bytes, dec._attn_ws_bytes)
        enc._attn_ws_bytes = dec._attn_ws_bytes = m
        P["enc"], P["dec"], P["cat"], P["skip_ch"] = enc.steps, dec.steps, cat, skip_ch
        return P

    # ------------------------------------------------------------------ execution
    def _stage(self, P, x, t, in_scale):
        assert x.shape[1:] == P["x"].shape[1:], "input shape %s does not match the model" % (tuple(x.shape),)
        P["x"].copy_(x)
        P["t"].copy_(t.reshape(-1).to(torch.float32))
        P["use_scale"][0] = in_scale is not None
        if in_scale is not None:
            P["in_scale"].copy_(in_scale.reshape(-1))

    def forward_scaled(self, x, noise_labels, in_scale=None):
        P = self._plan(x.shape[0])
        self._stage(P, x, noise_labels, in_scale)
        P["emb_n"][0] = self.emb_total
        run(P["enc"])
        run(P["dec"])
        return P["out"]

    def encode_scaled(self, x, noise_labels, in_scale=None):
        """NHWC fp32 [B,h,w,C] (plan buffer): the last encoder block's output."""
        P = self._plan(x.shape[0])
        self._stage(P, x, noise_labels, in_scale)
        P["emb_n"][0] = self.emb_enc
        run(P["enc"])
        return P["feat"]

    def __call__(self, x, noise_labels, class_labels=None, augment_labels=None):
        return self.forward_scaled(x, noise_labels).clone()

    forward = __call__

    def encode(self, x, noise_labels, class_labels=None, augment_labels=None):
        return self.encode_scaled(x, noise_labels).clone().permute(0, 3, 1, 2)

    def eval(self):
        return self

    def to(self, *a, **k):
        return self


class DhariwalUNet(SongUNet):
    """Drop-in for src/edm_networks.py:406 `DhariwalUNet` (inference, unconditional): the ADM architecture of the EDM code
    base - the same encoder / decoder skeleton as SongUNet with adaptive scale/shift blocks, 64-channel attention heads at
    every listed resolution, skip_scale 1, GroupNorm eps 1e-5, weight-less skips in the resampling blocks, the plain
    cos || sin embedding and `out_norm` / `out_conv`.  Like the reference class it has no `encode`: NLC 'pred*' styles need
    SongUNet; this is the base-style EDM model of BASELINE config 3."""
    _BLOCK_KW = dict(adaptive=True, head_ch=64, skip_scale=1.0, eps=1e-5)
    _EMB_ENDPOINT, _EMB_COS_FIRST = False, True
    _OUT_NORM, _OUT_CONV = "out_norm", "out_conv"
    _OUT_EPS = 1e-5

    def __init__(self, img_resolution, in_channels, out_channels, label_dim=0, augment_dim=0, model_channels=192,
                 channel_mult=(1, 2, 3, 4), channel_mult_emb=4, num_blocks=3, attn_resolutions=(32, 16, 8), dropout=0.10,
                 label_dropout=0, precision="bf16", device="cuda", **kwargs):
        for m in channel_mult:
            if (model_channels * m // 32) % 4 != 0 and model_channels * m >= 128:
                raise NotImplementedError("DhariwalUNet: %d channels give %d per GroupNorm group; the GroupNorm kernels take "
                                          "multiples of 4 (model_channels 128 / 256 work, 192 does not)"
                                          % (model_channels * m, model_channels * m // 32))
        super().__init__(img_resolution, in_channels, out_channels, label_dim=label_dim, augment_dim=augment_dim,
                         model_channels=model_channels, channel_mult=channel_mult, channel_mult_emb=channel_mult_emb,
                         num_blocks=num_blocks, attn_resolutions=attn_resolutions, precision=precision, device=device)

    @classmethod
    def from_reference(cls, m, precision="bf16", device="cuda"):
        raise NotImplementedError("construct DhariwalUNet with the reference's constructor arguments and load_state_dict()")

    def encode_scaled(self, x, noise_labels, in_scale=None):
        raise AttributeError("DhariwalUNet has no encode (src/edm_networks.py:406-502): use the 'base' styles")

    def encode(self, *a, **k):
        raise AttributeError("DhariwalUNet has no encode (src/edm_networks.py:406-502): use the 'base' styles")


class SigmaModel:
    """Drop-in for src/edm_networks.py:979 `SigmaModel`."""

    def __init__(self, dim=4, channels=64, n_blocks=2, out_dim=1, dropout=0.1, resample_filter=(1, 1),
                 precision="bf16", device="cuda"):
        if out_dim != 1:
            raise NotImplementedError("out_dim != 1 is not used by the reference")
        self.dim, self.channels, self.n_blocks = dim, channels, n_blocks
        d = dim
        for _ in range(n_blocks):
            if d % 2 != 0:
                raise NotImplementedError("odd feature sizes (ConstantPad2d branch, src/edm_networks.py:993-995) do "
                                          "not occur in the reference configurations")
            d //= 2
        self.final_dim = d
        self.eng = Engine(device, precision)
        self._plans, self._loaded = {}, False

    def load_state_dict(self, sd, strict=True):
        eng = self.eng
        sd = {k: v.detach() for k, v in sd.items()}
        C = self.channels
        self.blocks = []
        idx = 0
        for i in range(self.n_blocks):
            idx += 1
            blk = _BlockW(eng, sd, "down_layer.%d." % idx, with_emb=False)
            idx += 1
            p = "down_layer.%d.conv." % idx
            self.blocks.append((blk, eng.pack3x3(sd[p + "weight"]), eng.dev32(sd[p + "bias"])))
            idx += 1
        hw = self.final_dim * self.final_dim
        w = sd["fc_layer.1.weight"].float()
        w = w.view(-1, C, hw).permute(0, 2, 1).reshape(w.shape[0], hw * C)
        s = sd["fc_layer.2.weight"].float() / torch.sqrt(sd["fc_layer.2.running_var"].float() + 1e-5)
        self.fc_w = eng.dev32(w * s[:, None])
        self.fc_b = eng.dev32((sd["fc_layer.1.bias"].float() - sd["fc_layer.2.running_mean"].float()) * s
                              + sd["fc_layer.2.bias"].float())
        self.out_w, self.out_b = eng.dev32(sd["final_mlp.weight"]), eng.dev32(sd["final_mlp.bias"])
        self._loaded, self._plans = True, {}
        return self

    @classmethod
    def from_reference(cls, m, dim, precision="bf16", device="cuda"):
        blocks = [l for l in m.down_layer if type(l).__name__ == "PureUNetBlock"]
        self = cls(dim=dim, channels=blocks[0].in_channels, n_blocks=len(blocks), precision=precision, device=device)
        return self.load_state_dict(m.state_dict())

    def _plan(self, B):
        if B not in self._plans:
            self._plans[B] = self.eng.plan_two_pass(lambda: self._build_plan(B))
        return self._plans[B]

    def _build_plan(self, B):
        eng, C = self.eng, self.channels
        f32 = torch.float32
        P = {"feat": eng.named("sig.feat", (B, self.dim, self.dim, C), f32)}
        pc = PlanCtx(eng, B)
        cur = Feat(f32=Act(P["feat"]))
        res = self.dim
        for i, (blk, dw, db) in enumerate(self.blocks):
            o = Feat(f32=eng.act_f32("sg.o%d" % i, B, res, res, C), op=eng.act_op("sg.o16", B, res, res, C))
            _emit_block(pc, blk, cur, o)
            res //= 2
            d = Feat(f32=eng.act_f32("sg.d%d" % i, B, res, res, C))
            emit_conv3x3(pc, o.op, dw, db, C, d, stride=2, pad=0)  # pad (0,1,0,1) + stride-2 conv, :971-974
            cur = d
        flat = cur.f32.t.view(B, -1)
        P["hid"] = eng.named("sig.hid", (B, self.fc_w.shape[0]), f32)
        P["r"] = eng.named("sig.r", (B, 1), f32)
        pc.add(lambda: ops.linear(flat, self.fc_w, self.fc_b, P["hid"], act_out=1), "fc + BN + SiLU")
        pc.add(lambda: ops.linear(P["hid"], self.out_w, self.out_b, P["r"]), "final_mlp")
        P["steps"] = pc.steps
        return P

    def forward_nhwc(self, feat_nhwc):
        P = self._plan(feat_nhwc.shape[0])
        if feat_nhwc.data_ptr() != P["feat"].data_ptr():
            P["feat"].copy_(feat_nhwc)
        run(P["steps"])
        return P["r"]

    def __call__(self, feat):
        return self.forward_nhwc(feat.permute(0, 2, 3, 1)).clone().view(-1, 1, 1, 1)

    forward = __call__

    def eval(self):
        return self

    def to(self, *a, **k):
        return self
