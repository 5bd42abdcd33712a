"""Host-side mirror of the reference's ADM (guided-diffusion) UNet and its sigma-model (src/unet_adm.py),
executing on libnlc_b200 kernels.

`UNetModel` / `SigmaModel` take the reference constructors' arguments and consume the reference modules'
`state_dict()` unchanged.  Call conventions are the reference's: `model(x, t)` -> `[B, 3|6, R, R]`,
`model.encode(x, t)` -> `[B, Cf, hf, wf]` (input blocks + middle block, feat_layer = 1, src/unet_adm.py:668-693),
`sigma_model(feat)` -> `[B,1,1,1]`.  The sampler uses `forward_scaled` / `encode_scaled`, which fold the
per-sample input scale 1/sqrt(sigma^2+1) into the input convolution.

Layer mapping (src/unet_adm.py):
  ResBlock :143-256      GN32+SiLU -> [avg-pool / nearest-up of both h and x] -> conv3x3 -> GN32*(1+scale)+shift
                         (or +emb before the norm) -> SiLU -> conv3x3 (+ identity or fused 1x1 skip_connection)
  AttentionBlock :259-305  GN32 -> qkv GEMM -> softmax(q k^T / sqrt(ch)) v per head (legacy or new channel
                         order, :328-389) -> proj GEMM + x
  Downsample/Upsample :81-140 (conv_resample, used when resblock_updown is False)
"""
import math

import torch

from . import ops
from .engine import (Engine, Feat, PlanCtx, h_feat, emit_attention, emit_conv1x1, emit_conv3x3, emit_conv_in, emit_conv_out,
                     emit_groupnorm, run)
from .ops import Act

GN_EPS = 1e-5  # GroupNorm32 default, src/nn_util.py:93-100
GROUPS = 32


def _g(sd, key):
    if key not in sd:
        raise KeyError("state_dict is missing %r" % key)
    return sd[key]


class _ResW:
    """Packed weights of ResBlock / PureResNetBlock."""

    def __init__(self, eng, sd, p, scale_shift, updown=None, with_emb=True):
        g = lambda k: _g(sd, p + k)
        w1 = g("in_layers.2.weight")
        self.cin, self.cout = w1.shape[1], w1.shape[0]
        self.updown = updown
        self.scale_shift = scale_shift and with_emb
        self.n1w, self.n1b = eng.dev32(g("in_layers.0.weight")), eng.dev32(g("in_layers.0.bias"))
        self.n2w, self.n2b = eng.dev32(g("out_layers.0.weight")), eng.dev32(g("out_layers.0.bias"))
        self.w1, self.b1 = eng.pack3x3(w1), eng.dev32(g("in_layers.2.bias"))
        b2 = g("out_layers.3.bias").float()
        self.fused_skip = False
        if p + "skip_connection.weight" in sd:
            ws = g("skip_connection.weight")
            if ws.shape[-1] != 1:
                raise NotImplementedError("use_conv=True (3x3 skip_connection) is never built by the reference factories")
            self.w2 = eng.pack3x3(g("out_layers.3.weight"), extra=ws)
            b2 = b2 + g("skip_connection.bias").float()
            self.fused_skip = True
        else:
            self.w2 = eng.pack3x3(g("out_layers.3.weight"))
        self.b2 = eng.dev32(b2)
        self.emb_w = g("emb_layers.1.weight") if with_emb else None
        self.emb_b = g("emb_layers.1.bias") if with_emb else None
        self.emb_off = 0


class _AttnW:
    """AttentionBlock: qkv / proj_out are Conv1d(k=1) -> GEMMs."""

    def __init__(self, eng, sd, p, num_heads, num_head_channels, new_order):
        g = lambda k: _g(sd, p + k)
        self.C = g("proj_out.weight").shape[0]
        self.heads = num_heads if num_head_channels == -1 else self.C // num_head_channels
        self.dh = self.C // self.heads
        self.new_order = new_order
        self.nw, self.nb = eng.dev32(g("norm.weight")), eng.dev32(g("norm.bias"))
        self.wqkv = eng.pack3x3(g("qkv.weight").reshape(3 * self.C, self.C, 1, 1))
        self.bqkv = eng.dev32(g("qkv.bias"))
        self.wproj = eng.pack3x3(g("proj_out.weight").reshape(self.C, self.C, 1, 1))
        self.bproj = eng.dev32(g("proj_out.bias"))


def _emit_resblock(pc, w, x, dest, emb=None):
    """x: Feat (fp32, + operand copy when the block has a 1x1 skip); emb: [B, total] fused emb_layers output."""
    eng = pc.eng
    dt = eng.op_dtype
    B, H, W = x.B, x.H, x.W
    resid, resid_mode = x.res, 0
    if w.updown is None:
        a1 = eng.act_op("rb.a1", B, H, W, w.cin)
        emit_groupnorm(pc, x.res, w.n1w, w.n1b, GROUPS, GN_EPS, a1, silu=True)
    else:
        # h_upd / x_upd (src/unet_adm.py:236-243): the activated tensor is written already resampled by the
        # GroupNorm apply pass; x_upd is never materialised: the second conv's epilogue reads x at its own resolution
        # (nearest x2 / 2x2 average, nlc_conv_desc.resid_mode)
        mode = 1 if w.updown == "up" else 2
        src32 = x.res
        H, W = (2 * H, 2 * W) if mode == 1 else (H // 2, W // 2)
        a1 = eng.act_op("rb.a1r", B, H, W, w.cin)
        emit_groupnorm(pc, src32, w.n1w, w.n1b, GROUPS, GN_EPS, a1, silu=True, resample=mode)
        resid_mode = mode
    h = eng.act_h("rb.h", B, H, W, w.cout)
    rowvec = scale = shift = None
    if emb is not None and w.emb_w is not None:
        if w.scale_shift:
            scale = emb[:, w.emb_off:w.emb_off + w.cout]
            shift = emb[:, w.emb_off + w.cout:w.emb_off + 2 * w.cout]
        else:
            rowvec = emb[:, w.emb_off:w.emb_off + w.cout]
    emit_conv3x3(pc, a1, w.w1, w.b1, w.cout, h_feat(h), rowvec=rowvec)
    a2 = eng.act_op("rb.a2", B, H, W, w.cout)
    emit_groupnorm(pc, h, w.n2w, w.n2b, GROUPS, GN_EPS, a2, silu=True, scale=scale, shift=shift)
    if w.fused_skip:
        assert x.op is not None and w.updown is None
        emit_conv3x3(pc, a2, w.w2, w.b2, w.cout, dest, extra_src=x.op)
    else:
        emit_conv3x3(pc, a2, w.w2, w.b2, w.cout, dest, resid=resid, resid_mode=resid_mode)


def _emit_attnblock(pc, w, x, dest):
    eng = pc.eng
    B, H, W, C = x.B, x.H, x.W, w.C
    a = eng.act_op("at.a", B, H, W, C)
    emit_groupnorm(pc, x.res, w.nw, w.nb, GROUPS, GN_EPS, a, silu=False)
    qkv = eng.act_op("at.qkv", B, H, W, 3 * C)
    emit_conv1x1(pc, a, w.wqkv, w.bqkv, 3 * C, Feat(op=qkv))
    o = eng.act_op("at.o", B, H, W, C)
    scale = float(1.0 / math.sqrt(w.dh))  # (ch^-1/4)^2, src/unet_adm.py:346-349
    if w.new_order:  # [q | k | v] x [head][ch]
        emit_attention(pc, qkv, 0, C, 2 * C, w.dh, w.heads, w.dh, scale, o)
    else:  # legacy: [head] x [q | k | v][ch]
        emit_attention(pc, qkv, 0, w.dh, 2 * w.dh, 3 * w.dh, w.heads, w.dh, scale, o)
    emit_conv1x1(pc, o, w.wproj, w.bproj, C, dest, resid=x.res)


class UNetModel:
    """Drop-in for src/unet_adm.py:396 `UNetModel` (inference; unconditional, dims=2)."""

    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions,
                 dropout=0.0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2, num_classes=None,
                 use_checkpoint=False, use_fp16=False, num_heads=1, num_head_channels=-1, num_heads_upsample=-1,
                 use_scale_shift_norm=False, resblock_updown=False, use_new_attention_order=False, feat_layer=1,
                 precision="bf16", device="cuda"):
        if dims != 2 or num_classes is not None:
            raise NotImplementedError("only the unconditional 2-D model is on the sampling path (SURVEY §8a N1)")
        if not conv_resample and not resblock_updown:
            raise NotImplementedError("conv_resample=False without resblock_updown is not built by the factories")
        if feat_layer not in (0, 1):
            raise ValueError("feat_layer must be 0 or 1")
        self.image_size, self.in_channels, self.model_channels = image_size, in_channels, model_channels
        self.out_channels, self.num_res_blocks = out_channels, num_res_blocks
        self.attention_resolutions = tuple(attention_resolutions)
        self.channel_mult = tuple(channel_mult)
        self.num_heads, self.num_head_channels = num_heads, num_head_channels
        self.num_heads_upsample = num_heads if num_heads_upsample == -1 else num_heads_upsample
        self.use_scale_shift_norm, self.resblock_updown = use_scale_shift_norm, resblock_updown
        self.use_new_attention_order, self.feat_layer = use_new_attention_order, feat_layer
        self.emb_ch = 4 * model_channels
        self.eng = Engine(device, precision)
        self._plans, self._loaded = {}, False

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, sd, strict=True):
        eng = self.eng
        sd = {k: v.detach() for k, v in sd.items()}
        mc = self.model_channels
        ss, ru = self.use_scale_shift_norm, self.resblock_updown
        self.t0w, self.t0b = eng.dev32(sd["time_embed.0.weight"]), eng.dev32(sd["time_embed.0.bias"])
        self.t1w, self.t1b = eng.dev32(sd["time_embed.2.weight"]), eng.dev32(sd["time_embed.2.bias"])
        half = mc // 2
        # timestep_embedding (src/nn_util.py:113-116): exp(-log(1e4) * arange(half) / half), cos || sin
        self.freqs = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half).to(eng.device)
        self.cin_w, self.cin_b = eng.dev32(sd["input_blocks.0.0.weight"]), eng.dev32(sd["input_blocks.0.0.bias"])
        self.cin_wp = ops.pack_conv_in_weight(self.cin_w, eng.op_dtype)

        def attn(p, heads):
            return _AttnW(eng, sd, p, heads, self.num_head_channels, self.use_new_attention_order)

        def resampler(p, key):
            return (eng.pack3x3(sd[p + key + ".weight"]), eng.dev32(sd[p + key + ".bias"]))

        # each block: list of ("res", _ResW) | ("attn", _AttnW) | ("down", (w,b)) | ("up", (w,b))
        self.input_blocks, self.output_blocks = [], []
        ch = int(self.channel_mult[0] * mc)
        self.skip_ch = [ch]
        ds, idx = 1, 1
        L = len(self.channel_mult)
        for level, mult in enumerate(self.channel_mult):
            for _ in range(self.num_res_blocks):
                p = "input_blocks.%d." % idx
                layers = [("res", _ResW(eng, sd, p + "0.", ss))]
                ch = int(mult * mc)
                if ds in self.attention_resolutions:
                    layers.append(("attn", attn(p + "1.", self.num_heads)))
                self.input_blocks.append(layers)
                self.skip_ch.append(ch)
                idx += 1
            if level != L - 1:
                p = "input_blocks.%d.0." % idx
                if ru:
                    self.input_blocks.append([("res", _ResW(eng, sd, p, ss, updown="down"))])
                else:
                    self.input_blocks.append([("down", resampler(p, "op"))])
                self.skip_ch.append(ch)
                idx += 1
                ds *= 2
        self.mid_ch = ch
        self.middle = [("res", _ResW(eng, sd, "middle_block.0.", ss)), ("attn", attn("middle_block.1.", self.num_heads)),
                       ("res", _ResW(eng, sd, "middle_block.2.", ss))]
        idx = 0
        for level, mult in list(enumerate(self.channel_mult))[::-1]:
            for i in range(self.num_res_blocks + 1):
                p = "output_blocks.%d." % idx
                layers = [("res", _ResW(eng, sd, p + "0.", ss))]
                ch = int(mc * mult)
                j = 1
                if ds in self.attention_resolutions:
                    layers.append(("attn", attn(p + "%d." % j, self.num_heads_upsample)))
                    j += 1
                if level and i == self.num_res_blocks:
                    if ru:
                        layers.append(("res", _ResW(eng, sd, p + "%d." % j, ss, updown="up")))
                    else:
                        layers.append(("up", resampler(p + "%d." % j, "conv")))
                    ds //= 2
                self.output_blocks.append(layers)
                idx += 1
        self.no_w, self.no_b = eng.dev32(sd["out.0.weight"]), eng.dev32(sd["out.0.bias"])
        self.cout_w, self.cout_b = eng.dev32(sd["out.2.weight"]), eng.dev32(sd["out.2.bias"])
        self.cout_packed = (ops.pack_conv_out_weight(self.cout_w, self.cout_b, eng.op_dtype)
                            if eng.chunk == 64 and self.cout_w.shape[0] <= 8 else None)

        # one GEMM for every block's emb_layers; encoder + middle first so that encode() uses a prefix
        order = [l[1] for blk in self.input_blocks for l in blk if l[0] == "res"]
        order += [l[1] for l in self.middle if l[0] == "res"]
        n_enc = len(order)
        order += [l[1] for blk in self.output_blocks for l in blk if l[0] == "res"]
        off = 0
        for i, b in enumerate(order):
            if i == n_enc:
                self.emb_enc = off
            b.emb_off = off
            off += b.emb_w.shape[0]
        self.emb_total = off
        self.ew = eng.dev32(torch.cat([b.emb_w for b in order], dim=0))
        self.eb = eng.dev32(torch.cat([b.emb_b for b in order], dim=0))
        for b in order:
            b.emb_w = b.emb_b = True  # only the packed copy is kept
        self._loaded, self._plans = True, {}
        return self

    @classmethod
    def from_reference(cls, m, precision="bf16", device="cuda"):
        """Build from an instance of the reference's src.unet_adm.UNetModel."""
        ru = any(type(l).__name__ == "ResBlock" and l.updown for blk in m.input_blocks for l in blk)
        rb = next(l for blk in m.input_blocks for l in blk if type(l).__name__ == "ResBlock")
        at = m.middle_block[1]
        self = cls(m.image_size, m.in_channels, m.model_channels, m.out_channels, m.num_res_blocks,
                   m.attention_resolutions, channel_mult=m.channel_mult, num_heads=m.num_heads,
                   num_head_channels=m.num_head_channels, num_heads_upsample=m.num_heads_upsample,
                   use_scale_shift_norm=rb.use_scale_shift_norm, resblock_updown=ru,
                   use_new_attention_order=type(at.attention).__name__ == "QKVAttention",
                   feat_layer=m.feat_layer, precision=precision, device=device)
        return self.load_state_dict(m.state_dict())

    # ------------------------------------------------------------------ plan
    def _emit_block(self, pc, layers, x, dest, emb, B):
        """Run one TimestepEmbedSequential; the last layer writes `dest`."""
        eng = pc.eng
        cur = x
        for li, (kind, w) in enumerate(layers):
            last = li == len(layers) - 1
            if kind == "res":
                H, W = cur.H, cur.W
                if w.updown == "up":
                    H, W = 2 * H, 2 * W
                elif w.updown == "down":
                    H, W = H // 2, W // 2
                out = dest if last else eng.stream_feat("blk.o%d" % li, B, H, W, w.cout)
                _emit_resblock(pc, w, cur, out, emb)
            elif kind == "attn":
                out = dest if last else eng.stream_feat("blk.o%d" % li, B, cur.H, cur.W, w.C)
                _emit_attnblock(pc, w, cur, out)
            elif kind == "down":  # Downsample(use_conv): 3x3 stride 2 pad 1
                assert last
                dt = eng.op_dtype
                if cur.op is not None:
                    src = cur.op
                else:
                    src = eng.act_op("rs.src", B, cur.H, cur.W, cur.C)
                    c32 = cur.f32
                    pc.add(lambda c32=c32, src=src: ops.resample(c32, 0, None, src, dt))
                emit_conv3x3(pc, src, w[0], w[1], cur.C, dest, stride=2, pad=1)
                out = dest
            else:  # Upsample(use_conv): nearest x2 then 3x3
                assert last
                src = eng.act_op("rs.src", B, 2 * cur.H, 2 * cur.W, cur.C)
                c32 = cur.f32
                dt = eng.op_dtype
                if c32 is None:
                    pc.add(lambda c16=cur.op, src=src: ops.resample_op(c16, 1, src, dt))
                else:
                    pc.add(lambda c32=c32, src=src: ops.resample(c32, 1, None, src, dt))
                emit_conv3x3(pc, src, w[0], w[1], cur.C, dest)
                out = dest
            cur = out
        return cur

    def _plan(self, B):
        if B not in self._plans:
            assert self._loaded, "load_state_dict() first"
            self._plans[B] = self.eng.plan_two_pass(lambda: self._build_plan(B))
        return self._plans[B]

    def _build_plan(self, B):
        eng, R, mc = self.eng, self.image_size, self.model_channels
        f32, opt, dt = torch.float32, eng.op_torch, eng.op_dtype
        P = {}
        P["x"] = eng.named("x", (B, self.in_channels, R, R), f32)
        P["t"] = eng.named("t", (B,), f32)
        P["in_scale"] = eng.named("in_scale", (B,), f32)
        P["temb_sin"] = eng.named("temb_sin", (B, mc), f32)
        P["temb_h"] = eng.named("temb_h", (B, self.emb_ch), f32)
        P["temb"] = eng.named("temb", (B, self.emb_ch), f32)
        P["emb"] = eng.named("emb", (B, self.emb_total), f32)
        P["out"] = eng.named("out", (B, self.out_channels, R, R), f32)

        # resolution of every skip tensor (input conv, then each input block)
        res_list = [R]
        r = R
        for blk in self.input_blocks:
            if blk[0][0] == "down" or (blk[0][0] == "res" and blk[0][1].updown == "down"):
                r //= 2
            res_list.append(r)
        n_skips = len(self.skip_ch)
        # concat buffer for skip k: [h (c_h) | skip (c_skip)], consumed by output block n_skips-1-k
        cat = {}
        c_h = self.mid_ch
        for ob, k in zip(self.output_blocks, reversed(range(n_skips))):
            c_skip, rk = self.skip_ch[k], res_list[k]
            cat[k] = eng.cat_buffers(k, B, rk, rk, c_h + c_skip) + (c_h,)
            c_h = [l for l in ob if l[0] == "res"][-1][1].cout

        def skip_feat(k):
            c32, c16, c1 = cat[k]
            return eng.cat_view(c32, c16, c1, self.skip_ch[k])

        def head_feat(k):
            c32, c16, c1 = cat[k]
            return eng.cat_view(c32, c16, 0, c1)

        def cat_feat(k):
            c32, c16, _ = cat[k]
            return eng.cat_view(c32, c16)

        emb = P["emb"]
        enc = PlanCtx(eng, B)
        x_in, in_scale = P["x"], P["in_scale"]
        P["emb_n"] = [self.emb_enc]
        P["use_scale"] = [False]
        enc.add(lambda: ops.timestep_embedding(P["t"], self.freqs, True, P["temb_sin"]))
        enc.add(lambda: ops.linear(P["temb_sin"], self.t0w, self.t0b, P["temb_h"], act_out=1))
        enc.add(lambda: ops.linear(P["temb_h"], self.t1w, self.t1b, P["temb"]))
        enc.add(lambda: ops.linear(P["temb"], self.ew[:P["emb_n"][0]], self.eb[:P["emb_n"][0]],
                                   emb[:, :P["emb_n"][0]], act_in=1))
        d0 = skip_feat(0)
        emit_conv_in(enc, x_in, lambda: in_scale if P["use_scale"][0] else None, self.cin_wp, self.cin_b,
                     self.cin_w.shape[0], d0, w_f32=self.cin_w)
        cur = d0
        for k, blk in enumerate(self.input_blocks, start=1):
            cur = self._emit_block(enc, blk, cur, skip_feat(k), emb, B)
        rmid = res_list[-1]
        if self.feat_layer == 0:
            P["feat"] = None  # the last skip itself; copied out on demand
            P["feat_src"] = cur
        # middle block: the feature (feat_layer 1) is its output, which is also the decoder's first head
        mid_out = head_feat(n_skips - 1)
        feat32 = Act(eng.named("feat", (B, rmid, rmid, self.mid_ch), f32))
        mid = PlanCtx(eng, B)
        mid._gn_ws_floats, mid._attn_ws_bytes = enc._gn_ws_floats, enc._attn_ws_bytes
        # write the middle output into the concat head (fp32 + operand); the feature tensor is a dense copy
        if mid_out.f32 is None:
            # 16-bit residual stream at the middle resolution: the last conv writes the fp32 feature tensor itself next to
            # the operand copy in the concat head
            self._emit_block(mid, self.middle, cur, Feat(f32=feat32, op=mid_out.op), emb, B)
        else:
            self._emit_block(mid, self.middle, cur, mid_out, emb, B)
            src32 = mid_out.f32
            mid.add(lambda: ops.resample(src32, 0, feat32, None, dt))
        P["feat_mid"] = feat32.t

        dec = PlanCtx(eng, B)
        dec._gn_ws_floats, dec._attn_ws_bytes = mid._gn_ws_floats, mid._attn_ws_bytes
        k = n_skips - 1
        cur = None
        for ob in self.output_blocks:
            x = cat_feat(k)
            if k > 0:
                dest = head_feat(k - 1)
            else:
                dest = eng.stream_feat("up.out", B, R, R, [l for l in ob if l[0] == "res"][-1][1].cout)
            cur = self._emit_block(dec, ob, x, dest, emb, B)
            k -= 1
        a = eng.act_op("rb.a1", B, R, R, cur.C)
        emit_groupnorm(dec, cur.res, self.no_w, self.no_b, GROUPS, GN_EPS, a, silu=True)
        emit_conv_out(dec, a, self.cout_w, self.cout_b, self.cout_packed, P["out"])
        m = max(enc._gn_ws_floats, mid._gn_ws_floats, dec._gn_ws_floats)
        enc._gn_ws_floats = mid._gn_ws_floats = dec._gn_ws_floats = m
        m = max(enc._attn_ws_bytes, mid._attn_ws_bytes, dec._attn_ws_bytes)
        enc._attn_ws_bytes = mid._attn_ws_bytes = dec._attn_ws_bytes = m
        P["enc"], P["mid"], P["dec"] = enc.steps, mid.steps, dec.steps
        P["cat"] = cat  # skip k lives in channels [c1, c1+skip_ch[k]) of cat[k] (debug / tests)
        return P

    # ------------------------------------------------------------------ execution
    def _stage(self, P, x, t, in_scale):
        assert x.shape[1:] == P["x"].shape[1:], "input shape %s does not match the model" % (tuple(x.shape),)
        P["x"].copy_(x)
        P["t"].copy_(t.reshape(-1).to(torch.float32))
        P["use_scale"][0] = in_scale is not None
        if in_scale is not None:
            P["in_scale"].copy_(in_scale.reshape(-1))

    def forward_scaled(self, x, t, in_scale=None):
        P = self._plan(x.shape[0])
        self._stage(P, x, t, in_scale)
        P["emb_n"][0] = self.emb_total
        run(P["enc"])
        run(P["mid"])
        run(P["dec"])
        return P["out"]

    def encode_scaled(self, x, t, in_scale=None):
        """NHWC fp32 [B,h,w,C] feature (plan buffer): middle-block output (feat_layer 1) or the last input block's
        output (feat_layer 0)."""
        P = self._plan(x.shape[0])
        self._stage(P, x, t, in_scale)
        P["emb_n"][0] = self.emb_enc
        run(P["enc"])
        if self.feat_layer == 0:
            return P["feat_src"].res.dense().float().contiguous()
        run(P["mid"])
        return P["feat_mid"]

    def __call__(self, x, t, y=None):
        return self.forward_scaled(x, t).clone()

    forward = __call__

    def encode(self, x, t, y=None):
        return self.encode_scaled(x, t).clone().permute(0, 3, 1, 2)

    def forward_and_encode(self, x, t, y=None):
        out = self.forward_scaled(x, t).clone()
        P = self._plan(x.shape[0])
        feat = P["feat_src"].res.dense().float() if self.feat_layer == 0 else P["feat_mid"]
        return out, feat.clone().permute(0, 3, 1, 2)

    def convert_to_fp16(self):
        """The reference's reduced-precision switch (src/unet_adm.py:619-625); precision is chosen at construction
        here (bf16 / tf32 operands), so this is a no-op kept for API compatibility."""
        return self

    convert_to_fp32 = convert_to_fp16

    def eval(self):
        return self

    def to(self, *a, **k):
        return self


class SigmaModel:
    """Drop-in for src/unet_adm.py:1029 `SigmaModel`."""

    def __init__(self, dim=4, channels=64, n_blocks=2, out_dim=1, dropout=0.1, num_heads=1, num_head_channels=-1,
                 use_new_attention_order=False, use_checkpoint=False, use_fp16=False, precision="bf16", device="cuda"):
        if out_dim != 1:
            raise NotImplementedError("out_dim != 1 is not used by the reference")
        self.dim, self.channels, self.n_blocks = dim, channels, n_blocks
        self.num_heads, self.num_head_channels, self.new_order = num_heads, num_head_channels, use_new_attention_order
        d = dim
        for _ in range(n_blocks):
            if d % 2 != 0:
                raise NotImplementedError("odd feature sizes (ConstantPad2d branch, src/unet_adm.py:1038-1040) do "
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
            blk = {"res": _ResW(eng, sd, "down_layer.%d." % idx, False, with_emb=False)}
            idx += 1
            if i == 0:
                blk["attn"] = _AttnW(eng, sd, "down_layer.%d." % idx, self.num_heads, self.num_head_channels,
                                     self.new_order)
                idx += 1
            p = "down_layer.%d.op." % idx
            blk["down"] = (eng.pack3x3(sd[p + "weight"]), eng.dev32(sd[p + "bias"]))
            idx += 1
            self.blocks.append(blk)
        hw = self.final_dim * self.final_dim
        w = sd["fc_layer.1.weight"].float()
        w = w.view(-1, C, hw).permute(0, 2, 1).reshape(w.shape[0], hw * C)  # NCHW-flatten -> NHWC-flatten columns
        s = sd["fc_layer.2.weight"].float() / torch.sqrt(sd["fc_layer.2.running_var"].float() + 1e-5)
        self.fc_w = eng.dev32(w * s[:, None])
        self.fc_b = eng.dev32((sd["fc_layer.1.bias"].float() - sd["fc_layer.2.running_mean"].float()) * s
                              + sd["fc_layer.2.bias"].float())
        self.out_w, self.out_b = eng.dev32(sd["final_mlp.weight"]), eng.dev32(sd["final_mlp.bias"])
        self._loaded, self._plans = True, {}
        return self

    @classmethod
    def from_reference(cls, m, dim, precision="bf16", device="cuda"):
        res = [l for l in m.down_layer if type(l).__name__ == "PureResNetBlock"]
        at = next(l for l in m.down_layer if type(l).__name__ == "AttentionBlock")
        self = cls(dim=dim, channels=res[0].channels, n_blocks=len(res), num_heads=at.num_heads,
                   use_new_attention_order=type(at.attention).__name__ == "QKVAttention", precision=precision,
                   device=device)
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
        for i, blk in enumerate(self.blocks):
            has_attn = "attn" in blk
            o = Feat(f32=eng.act_f32("sg.o%d" % i, B, res, res, C),
                     op=None if has_attn else eng.act_op("sg.o16", B, res, res, C))
            _emit_resblock(pc, blk["res"], cur, o)
            if has_attn:
                o2 = Feat(f32=eng.act_f32("sg.p%d" % i, B, res, res, C), op=eng.act_op("sg.o16", B, res, res, C))
                _emit_attnblock(pc, blk["attn"], o, o2)
                o = o2
            res //= 2
            d = Feat(f32=eng.act_f32("sg.d%d" % i, B, res, res, C))
            emit_conv3x3(pc, o.op, blk["down"][0], blk["down"][1], C, d, stride=2, pad=1)
            cur = d
        flat = cur.f32.t.view(B, -1)
        P["hid"] = eng.named("sig.hid", (B, self.fc_w.shape[0]), f32)
        P["r"] = eng.named("sig.r", (B, 1), f32)
        pc.add(lambda: ops.linear(flat, self.fc_w, self.fc_b, P["hid"], act_out=2))
        pc.add(lambda: ops.linear(P["hid"], self.out_w, self.out_b, P["r"]))
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

    def convert_to_fp16(self):
        return self

    convert_to_fp32 = convert_to_fp16

    def eval(self):
        return self

    def to(self, *a, **k):
        return self
