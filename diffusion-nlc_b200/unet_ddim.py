"""Host-side mirror of the reference's DDPM/DDIM UNet and its sigma-model (src/unet_ddim.py), executing on
libnlc_b200 kernels.

`UNetModel` and `SigmaModel` take the reference constructors' arguments and consume the reference modules'
`state_dict()` unchanged (same keys), so `create_simple_sigma_eps_model` checkpoints drop in.  Calls keep the
reference conventions: `model(x, t)` / `model.encode(x, t)` with x `[B,3,R,R]` fp32 NCHW (already scaled by
1/sqrt(sigma^2+1), src/experiments.py:295-311) and t `[B]`; `sigma_model(feat)` -> `[B,1,1,1]`.  Internally
everything is NHWC; the fused entry points used by the sampler (`forward_scaled`, `encode_scaled`) additionally
fold the per-sample input scale into conv_in.
"""
import math
import os

import torch

from . import ops
from .engine import (Engine, Feat, PlanCtx, h_feat, emit_attention, emit_conv1x1, emit_conv3x3, emit_conv_in, emit_conv_out,
                     emit_groupnorm, emit_upsample_conv3x3, run, upsample_conv_eligible)
from .ops import Act

GN_EPS = 1e-6  # Normalize(), src/unet_ddim.py:54-55
GROUPS = 32


def _sd_get(sd, key):
    if key not in sd:
        raise KeyError("state_dict is missing %r" % key)
    return sd[key]


class _ResBlockW:
    """Packed weights of ResnetBlock / PureResnetBlock (src/unet_ddim.py:99-156, 438-490)."""

    def __init__(self, eng, sd, prefix, with_temb=True):
        g = lambda k: _sd_get(sd, prefix + k)
        self.cin = g("conv1.weight").shape[1]
        self.cout = g("conv1.weight").shape[0]
        self.n1w, self.n1b = eng.dev32(g("norm1.weight")), eng.dev32(g("norm1.bias"))
        self.n2w, self.n2b = eng.dev32(g("norm2.weight")), eng.dev32(g("norm2.bias"))
        self.w1, self.b1 = eng.pack3x3(g("conv1.weight")), eng.dev32(g("conv1.bias"))
        # with a time embedding the conv1 bias rides on the per-sample temb row (UNetModel.load_state_dict adds it to the
        # fused temb_proj bias): one vector instead of two in the conv epilogue
        self.b1_host = g("conv1.bias").float().cpu() if with_temb else None
        self.shortcut = None
        b2 = g("conv2.bias").float()
        if prefix + "nin_shortcut.weight" in sd:
            self.shortcut = "nin"
            self.w2 = eng.pack3x3(g("conv2.weight"), extra=g("nin_shortcut.weight"))
            b2 = b2 + g("nin_shortcut.bias").float()
        elif prefix + "conv_shortcut.weight" in sd:
            raise NotImplementedError("conv_shortcut=True is never instantiated by the reference factories")
        else:
            self.w2 = eng.pack3x3(g("conv2.weight"))
        self.b2 = eng.dev32(b2)
        self.temb_w = g("temb_proj.weight") if with_temb else None
        self.temb_b = g("temb_proj.bias") if with_temb else None


class _AttnW:
    """AttnBlock (src/unet_ddim.py:159-211): q,k,v 1x1 convs concatenated into one GEMM."""

    def __init__(self, eng, sd, prefix):
        g = lambda k: _sd_get(sd, prefix + k)
        self.C = g("q.weight").shape[0]
        self.nw, self.nb = eng.dev32(g("norm.weight")), eng.dev32(g("norm.bias"))
        wqkv = torch.cat([g("q.weight"), g("k.weight"), g("v.weight")], dim=0)
        self.wqkv = eng.pack3x3(wqkv)
        self.bqkv = eng.dev32(torch.cat([g("q.bias"), g("k.bias"), g("v.bias")]))
        self.wproj = eng.pack3x3(g("proj_out.weight"))
        self.bproj = eng.dev32(g("proj_out.bias"))


def _emit_resblock(pc, wts, x, dest, rowvec=None):
    """GN+SiLU -> conv3x3 (+bias +temb) -> GN+SiLU -> conv3x3 (+bias, + identity or fused 1x1 shortcut)."""
    eng = pc.eng
    B, H, W = x.B, x.H, x.W
    a1 = eng.act_op("rb.a1", B, H, W, wts.cin)
    emit_groupnorm(pc, x.res, wts.n1w, wts.n1b, GROUPS, GN_EPS, a1, silu=True)
    h = eng.act_h("rb.h", B, H, W, wts.cout)
    emit_conv3x3(pc, a1, wts.w1, wts.b1 if rowvec is None else None, wts.cout, h_feat(h), rowvec=rowvec)
    a2 = eng.act_op("rb.a2", B, H, W, wts.cout)
    emit_groupnorm(pc, h, wts.n2w, wts.n2b, GROUPS, GN_EPS, a2, silu=True)
    if wts.shortcut == "nin":
        assert x.op is not None, "a block with a 1x1 shortcut needs the operand copy of its input"
        emit_conv3x3(pc, a2, wts.w2, wts.b2, wts.cout, dest, extra_src=x.op)
    else:
        emit_conv3x3(pc, a2, wts.w2, wts.b2, wts.cout, dest, resid=x.res)


def _emit_attnblock(pc, wts, x, dest):
    eng = pc.eng
    B, H, W, C = x.B, x.H, x.W, wts.C
    a = eng.act_op("at.a", B, H, W, C)
    emit_groupnorm(pc, x.res, wts.nw, wts.nb, GROUPS, GN_EPS, a, silu=False)
    qkv = eng.act_op("at.qkv", B, H, W, 3 * C)
    emit_conv1x1(pc, a, wts.wqkv, wts.bqkv, 3 * C, Feat(op=qkv))
    o = eng.act_op("at.o", B, H, W, C)
    emit_attention(pc, qkv, 0, C, 2 * C, 0, 1, C, float(int(C) ** (-0.5)), o)
    emit_conv1x1(pc, o, wts.wproj, wts.bproj, C, dest, resid=x.res)


def _emit_downsample(pc, w, b, x, dest):
    """Downsample(with_conv): pad (0,1,0,1) then 3x3 stride 2 (src/unet_ddim.py:89-94)."""
    assert x.op is not None
    emit_conv3x3(pc, x.op, w, b, x.C, dest, stride=2, pad=0)


class UNetModel:
    """Drop-in for src/unet_ddim.py:214 `UNetModel` (inference only)."""

    def __init__(self, image_size, in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions,
                 dropout=0.0, channel_mult=(1, 2, 4, 8), conv_resample=True, feat_layer=0, precision="bf16", device="cuda"):
        # feat_layer: which middle-block tensor `encode` returns - 0: after mid.attn_1 (src/unet_ddim.py:365-393 and
        # src/unet_simple.py:371-372), 1: after mid.block_2 (src/unet_simple.py:373-375,402-407: `config.model.feat_layer`)
        if feat_layer not in (0, 1):
            raise ValueError("feat_layer must be 0 or 1")
        self.feat_layer = feat_layer
        if not conv_resample:
            raise NotImplementedError("conv_resample=False (avg-pool / bare upsample) is not used by the reference "
                                      "factories (src/script_util.py:209-219)")
        self.resolution = image_size
        self.in_channels = in_channels
        self.ch = model_channels
        self.temb_ch = 4 * model_channels
        self.out_ch = out_channels
        self.num_res_blocks = num_res_blocks
        self.attn_resolutions = tuple(attention_resolutions)
        self.ch_mult = tuple(channel_mult)
        self.num_resolutions = len(self.ch_mult)
        self.eng = Engine(device, precision)
        self._plans = {}
        self._loaded = False

    # ------------------------------------------------------------------ weights
    def load_state_dict(self, sd, strict=True):
        eng = self.eng
        sd = {k: v.detach() for k, v in sd.items()}
        self.t0w, self.t0b = eng.dev32(sd["temb.dense.0.weight"]), eng.dev32(sd["temb.dense.0.bias"])
        self.t1w, self.t1b = eng.dev32(sd["temb.dense.1.weight"]), eng.dev32(sd["temb.dense.1.bias"])
        self.cin_w, self.cin_b = eng.dev32(sd["conv_in.weight"]), eng.dev32(sd["conv_in.bias"])
        self.cin_wp = ops.pack_conv_in_weight(self.cin_w, eng.op_dtype)
        half = self.ch // 2
        # get_timestep_embedding (src/unet_ddim.py:38-41): exp(arange(half) * -(log(1e4)/(half-1))) in fp32
        emb = math.log(10000) / (half - 1)
        self.freqs = torch.exp(torch.arange(half, dtype=torch.float32) * -emb).to(eng.device)

        self.down, self.up = [], []
        curr_res = self.resolution
        for i_level in range(self.num_resolutions):
            lvl = {"block": [], "attn": [], "down": None}
            for i_block in range(self.num_res_blocks):
                lvl["block"].append(_ResBlockW(eng, sd, "down.%d.block.%d." % (i_level, i_block)))
                if curr_res in self.attn_resolutions:
                    lvl["attn"].append(_AttnW(eng, sd, "down.%d.attn.%d." % (i_level, i_block)))
            if i_level != self.num_resolutions - 1:
                p = "down.%d.downsample.conv." % i_level
                lvl["down"] = (eng.pack3x3(sd[p + "weight"]), eng.dev32(sd[p + "bias"]))
                curr_res //= 2
            self.down.append(lvl)
        self.mid1 = _ResBlockW(eng, sd, "mid.block_1.")
        self.mid_attn = _AttnW(eng, sd, "mid.attn_1.")
        self.mid2 = _ResBlockW(eng, sd, "mid.block_2.")
        for i_level in range(self.num_resolutions):
            lvl = {"block": [], "attn": [], "up": None}
            res = self.resolution // (2 ** i_level)
            for i_block in range(self.num_res_blocks + 1):
                lvl["block"].append(_ResBlockW(eng, sd, "up.%d.block.%d." % (i_level, i_block)))
                if res in self.attn_resolutions:
                    lvl["attn"].append(_AttnW(eng, sd, "up.%d.attn.%d." % (i_level, i_block)))
            if i_level != 0:
                p = "up.%d.upsample.conv." % i_level
                lvl["up"] = (eng.pack3x3(sd[p + "weight"]), eng.dev32(sd[p + "bias"]))
                lvl["up_phase"] = ops.upsample_phase_weights(sd[p + "weight"].to(eng.device), eng.op_dtype)
            self.up.append(lvl)
        self.no_w, self.no_b = eng.dev32(sd["norm_out.weight"]), eng.dev32(sd["norm_out.bias"])
        self.cout_w, self.cout_b = eng.dev32(sd["conv_out.weight"]), eng.dev32(sd["conv_out.bias"])
        self.cout_packed = (ops.pack_conv_out_weight(self.cout_w, self.cout_b, eng.op_dtype)
                            if eng.chunk == 64 and self.cout_w.shape[0] <= 8 else None)

        # one GEMM for every block's temb projection; encoder blocks first so encode() uses a prefix
        order = [b for lvl in self.down for b in lvl["block"]] + [self.mid1, self.mid2]
        order += [b for lvl in reversed(self.up) for b in lvl["block"]]
        off = 0
        for b in order:
            b.temb_off = off
            off += b.cout
        self.temb_total = off
        self.temb_enc = (self.mid2 if self.feat_layer else self.mid1).temb_off + (self.mid2 if self.feat_layer else self.mid1).cout
        self.tpw = eng.dev32(torch.cat([b.temb_w for b in order], dim=0))
        self.tpb = eng.dev32(torch.cat([b.temb_b.float().cpu() + b.b1_host for b in order], dim=0))
        self._loaded = True
        self._plans = {}
        return self

    @classmethod
    def from_reference(cls, ref_module, precision="bf16", device="cuda"):
        """Build from an instance of the reference's src.unet_ddim.UNetModel or src.unet_simple.Model (either feat_layer)."""
        m = ref_module
        attn_res = sorted({m.resolution // (2 ** i) for i, lvl in enumerate(m.down) if len(lvl.attn) > 0})
        ch_mult = tuple(lvl.block[-1].out_channels // m.ch for lvl in m.down)
        self = cls(m.resolution, m.in_channels, m.ch, m.conv_out.out_channels, m.num_res_blocks, attn_res,
                   channel_mult=ch_mult, feat_layer=int(getattr(m, "feat_layer", 0) != 0), precision=precision, device=device)
        return self.load_state_dict(m.state_dict())

    # ------------------------------------------------------------------ plan
    def _plan(self, B):
        if B not in self._plans:
            assert self._loaded, "load_state_dict() first"
            self._plans[B] = self.eng.plan_two_pass(lambda: self._build_plan(B))
        return self._plans[B]

    def _build_plan(self, B):
        eng, R, ch = self.eng, self.resolution, self.ch
        f32, opt = torch.float32, eng.op_torch
        P = {}
        # I/O buffers
        P["x"] = eng.named("x", (B, self.in_channels, R, R), f32)
        P["t"] = eng.named("t", (B,), f32)
        P["in_scale"] = eng.named("in_scale", (B,), f32)
        P["temb_sin"] = eng.named("temb_sin", (B, ch), f32)
        P["temb_h"] = eng.named("temb_h", (B, self.temb_ch), f32)
        P["temb"] = eng.named("temb", (B, self.temb_ch), f32)
        P["tp"] = eng.named("tp", (B, self.temb_total), f32)
        P["out"] = eng.named("out", (B, self.out_ch, R, R), f32)

        # ---- shapes of the skip stack, then the concat buffer each skip lands in
        hs_shapes = [(ch, R)]
        res = R
        for i_level in range(self.num_resolutions):
            for _ in range(self.num_res_blocks):
                hs_shapes.append((ch * self.ch_mult[i_level], res))
            if i_level != self.num_resolutions - 1:
                res //= 2
                hs_shapes.append((ch * self.ch_mult[i_level], res))
        c_mid = ch * self.ch_mult[-1]
        cat = {}  # skip index -> (cat32, cat16, C1)
        k = len(hs_shapes) - 1
        c_h = c_mid
        for i_level in reversed(range(self.num_resolutions)):
            for i_block in range(self.num_res_blocks + 1):
                c_skip, r = hs_shapes[k]
                cat[k] = eng.cat_buffers(k, B, r, r, c_h + c_skip) + (c_h,)
                c_h = self.up[i_level]["block"][i_block].cout
                k -= 1

        def skip_feat(k):
            c32, c16, c1 = cat[k]
            return eng.cat_view(c32, c16, c1, hs_shapes[k][0])

        def head_feat(k):
            c32, c16, c1 = cat[k]
            return eng.cat_view(c32, c16, 0, c1)

        def cat_feat(k):
            c32, c16, _ = cat[k]
            return eng.cat_view(c32, c16)

        # ---- encoder (shared by forward and encode)
        enc = PlanCtx(eng, B)
        tp = P["tp"]
        x_in, in_scale = P["x"], P["in_scale"]
        d0 = skip_feat(0)
        dt = eng.op_dtype
        enc.add(lambda: ops.timestep_embedding(P["t"], self.freqs, False, P["temb_sin"]))
        enc.add(lambda: ops.linear(P["temb_sin"], self.t0w, self.t0b, P["temb_h"], act_out=1))
        enc.add(lambda: ops.linear(P["temb_h"], self.t1w, self.t1b, P["temb"]))
        P["tp_n"] = [self.temb_enc]  # mutable: forward sets temb_total, encode temb_enc
        enc.add(lambda: ops.linear(P["temb"], self.tpw[:P["tp_n"][0]], self.tpb[:P["tp_n"][0]],
                                   tp[:, :P["tp_n"][0]], act_in=1))
        P["use_scale"] = [False]
        emit_conv_in(enc, x_in, lambda: in_scale if P["use_scale"][0] else None, self.cin_wp, self.cin_b,
                     self.cin_w.shape[0], d0, w_f32=self.cin_w)
        k = 0
        cur = d0
        res = R
        for i_level in range(self.num_resolutions):
            lvl = self.down[i_level]
            for i_block in range(self.num_res_blocks):
                wts = lvl["block"][i_block]
                k += 1
                dest = skip_feat(k)
                rv = tp[:, wts.temb_off:wts.temb_off + wts.cout]
                if lvl["attn"]:
                    tmp = eng.stream_feat("blk.out", B, res, res, wts.cout)
                    _emit_resblock(enc, wts, cur, tmp, rowvec=rv)
                    _emit_attnblock(enc, lvl["attn"][i_block], tmp, dest)
                else:
                    _emit_resblock(enc, wts, cur, dest, rowvec=rv)
                cur = dest
            if lvl["down"] is not None:
                k += 1
                dest = skip_feat(k)
                _emit_downsample(enc, lvl["down"][0], lvl["down"][1], cur, dest)
                cur = dest
                res //= 2
        n_skips = k
        rmid = res
        m1 = eng.stream_feat("mid.1", B, rmid, rmid, c_mid)
        _emit_resblock(enc, self.mid1, cur, m1, rowvec=tp[:, self.mid1.temb_off:self.mid1.temb_off + c_mid])
        feat = Feat(f32=Act(eng.named("feat", (B, rmid, rmid, c_mid), f32)))
        P["feat"] = feat.f32.t
        k = n_skips
        rv2 = tp[:, self.mid2.temb_off:self.mid2.temb_off + c_mid]
        dec = PlanCtx(eng, B)
        if self.feat_layer == 0:
            _emit_attnblock(enc, self.mid_attn, m1, feat)
        else:
            # the feature is mid.block_2's output, which is also the decoder's first head: the block belongs to the encoder
            # pass; it writes the head slice of the first concat buffer and (fp32) the feature tensor
            ma = eng.stream_feat("mid.a", B, rmid, rmid, c_mid)
            _emit_attnblock(enc, self.mid_attn, m1, ma)
            head = head_feat(k)
            if head.f32 is None:
                _emit_resblock(enc, self.mid2, ma, Feat(f32=feat.f32, op=head.op), rowvec=rv2)
            else:
                _emit_resblock(enc, self.mid2, ma, head, rowvec=rv2)
                h32, f32t = head.f32, feat.f32
                enc.add(lambda: ops.resample(h32, 0, f32t, None, dt), "feat copy")

        # ---- decoder
        dec._gn_ws_floats, dec._attn_ws_bytes = enc._gn_ws_floats, enc._attn_ws_bytes
        if self.feat_layer == 0:
            _emit_resblock(dec, self.mid2, feat, head_feat(k), rowvec=rv2)
        res = rmid
        for i_level in reversed(range(self.num_resolutions)):
            lvl = self.up[i_level]
            for i_block in range(self.num_res_blocks + 1):
                wts = lvl["block"][i_block]
                x = cat_feat(k)
                last_in_level = i_block == self.num_res_blocks
                last = last_in_level and i_level == 0
                if last or last_in_level:
                    dest = eng.stream_feat("up.out", B, res, res, wts.cout)
                else:
                    dest = head_feat(k - 1)
                rv = tp[:, wts.temb_off:wts.temb_off + wts.cout]
                if lvl["attn"]:
                    tmp = eng.stream_feat("blk.out", B, res, res, wts.cout)
                    _emit_resblock(dec, wts, x, tmp, rowvec=rv)
                    _emit_attnblock(dec, lvl["attn"][i_block], tmp, dest)
                else:
                    _emit_resblock(dec, wts, x, dest, rowvec=rv)
                cur = dest
                k -= 1
            if lvl["up"] is not None:
                # Upsample: nearest x2 then 3x3 conv (src/unet_ddim.py:69-74); the replicated operand is
                # materialised once in the operand dtype
                src32 = cur.f32  # (None on the 16-bit residual stream: the block output already is the operand)
                if upsample_conv_eligible(res, res):
                    # ... computed at the low resolution instead: four sub-pixel phase convs (engine.emit_upsample_conv3x3)
                    if src32 is None:
                        lowo = cur.op
                    else:
                        lowo = eng.act_op("up.low", B, res, res, cur.C)
                        dec.add(lambda src32=src32, lowo=lowo: ops.resample(src32, 0, None, lowo, dt))
                    res *= 2
                    emit_upsample_conv3x3(dec, lowo, lvl["up_phase"], lvl["up"][1], cur.C, head_feat(k))
                else:
                    upo = eng.act_op("up.rep", B, 2 * res, 2 * res, cur.C)
                    if src32 is None:
                        dec.add(lambda src=cur.op, upo=upo: ops.resample_op(src, 1, upo, dt))
                    else:
                        dec.add(lambda src32=src32, upo=upo: ops.resample(src32, 1, None, upo, dt))
                    res *= 2
                    emit_conv3x3(dec, upo, lvl["up"][0], lvl["up"][1], cur.C, head_feat(k))
        a = eng.act_op("rb.a1", B, R, R, cur.C)
        emit_groupnorm(dec, cur.res, self.no_w, self.no_b, GROUPS, GN_EPS, a, silu=True)
        emit_conv_out(dec, a, self.cout_w, self.cout_b, self.cout_packed, P["out"])
        enc._gn_ws_floats = dec._gn_ws_floats = max(enc._gn_ws_floats, dec._gn_ws_floats)
        enc._attn_ws_bytes = dec._attn_ws_bytes = max(enc._attn_ws_bytes, dec._attn_ws_bytes)
        P["enc"], P["dec"] = enc.steps, dec.steps
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
        """eps (and learned-variance channels) for x*in_scale[b]; returns the plan's NCHW output buffer."""
        P = self._plan(x.shape[0])
        self._stage(P, x, t, in_scale)
        P["tp_n"][0] = self.temb_total
        run(P["enc"])
        run(P["dec"])
        return P["out"]

    def encode_scaled(self, x, t, in_scale=None):
        """Feature after mid.attn_1 (src/unet_ddim.py:365-393) as NHWC fp32 [B,h,w,C] (plan buffer)."""
        P = self._plan(x.shape[0])
        self._stage(P, x, t, in_scale)
        P["tp_n"][0] = self.temb_enc
        run(P["enc"])
        return P["feat"]

    def __call__(self, x, t):
        return self.forward_scaled(x, t).clone()

    forward = __call__

    def encode(self, x, t):
        """Reference layout: [B,C,h,w] (a permuted view of an NHWC tensor)."""
        return self.encode_scaled(x, t).clone().permute(0, 3, 1, 2)

    def forward_and_encode(self, x, t):
        out = self.forward_scaled(x, t).clone()
        P = self._plan(x.shape[0])
        return out, P["feat"].clone().permute(0, 3, 1, 2)

    def eval(self):
        return self

    def to(self, *a, **k):
        return self


class SigmaModel:
    """Drop-in for src/unet_ddim.py:493 `SigmaModel`: the noise-level-correction head r_hat(feat)."""

    def __init__(self, dim=4, channels=64, n_blocks=2, out_dim=1, dropout=0.1, precision="bf16", device="cuda"):
        if out_dim != 1:
            raise NotImplementedError("out_dim != 1 is not used by the reference")
        self.dim, self.channels, self.n_blocks = dim, channels, n_blocks
        d = dim
        for _ in range(n_blocks):
            if d % 2 != 0:
                raise NotImplementedError("odd feature sizes (ConstantPad2d branch, src/unet_ddim.py:499-501) "
                                          "do not occur in the reference configurations")
            d //= 2
        self.final_dim = d
        self.eng = Engine(device, precision)
        self._plans = {}
        self._loaded = False

    def load_state_dict(self, sd, strict=True):
        eng = self.eng
        sd = {k: v.detach() for k, v in sd.items()}
        C = self.channels
        self.blocks = []
        idx = 0
        for i in range(self.n_blocks):
            idx += 1  # Identity / pad slot
            blk = {"res": _ResBlockW(eng, sd, "down_layer.%d." % idx, with_temb=False)}
            idx += 1
            if i == 0:
                blk["attn"] = _AttnW(eng, sd, "down_layer.%d." % idx)
                idx += 1
            p = "down_layer.%d.conv." % idx
            blk["down"] = (eng.pack3x3(sd[p + "weight"]), eng.dev32(sd[p + "bias"]))
            idx += 1
            self.blocks.append(blk)
        # Flatten (NCHW order) -> Linear -> BatchNorm1d(eval) -> GELU -> Linear  (src/unet_ddim.py:514-519)
        hw = self.final_dim * self.final_dim
        w = sd["fc_layer.1.weight"].float()  # [128, C*hw], column index c*hw + p
        w = w.view(-1, C, hw).permute(0, 2, 1).reshape(w.shape[0], hw * C)  # -> column index p*C + c (NHWC)
        s = sd["fc_layer.2.weight"].float() / torch.sqrt(sd["fc_layer.2.running_var"].float() + 1e-5)
        self.fc_w = eng.dev32(w * s[:, None])
        self.fc_b = eng.dev32((sd["fc_layer.1.bias"].float() - sd["fc_layer.2.running_mean"].float()) * s
                              + sd["fc_layer.2.bias"].float())
        self.out_w, self.out_b = eng.dev32(sd["final_mlp.weight"]), eng.dev32(sd["final_mlp.bias"])
        self._loaded = True
        self._plans = {}
        return self

    @classmethod
    def from_reference(cls, ref_module, dim, precision="bf16", device="cuda"):
        m = ref_module
        channels = m.down_layer[1].in_channels
        n_blocks = sum(1 for l in m.down_layer if type(l).__name__ == "Downsample")
        self = cls(dim=dim, channels=channels, n_blocks=n_blocks, precision=precision, device=device)
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
            _emit_downsample(pc, blk["down"][0], blk["down"][1], o, d)
            cur = d
        flat = cur.f32.t.view(B, -1)
        P["hid"] = eng.named("sig.hid", (B, self.fc_w.shape[0]), f32)
        P["r"] = eng.named("sig.r", (B, 1), f32)
        pc.add(lambda: ops.linear(flat, self.fc_w, self.fc_b, P["hid"], act_out=2))
        pc.add(lambda: ops.linear(P["hid"], self.out_w, self.out_b, P["r"]))
        P["steps"] = pc.steps
        return P

    def forward_nhwc(self, feat_nhwc):
        """feat [B,h,w,C] fp32 contiguous -> r [B,1] (plan buffer)."""
        P = self._plan(feat_nhwc.shape[0])
        if feat_nhwc.data_ptr() != P["feat"].data_ptr():
            P["feat"].copy_(feat_nhwc)
        run(P["steps"])
        return P["r"]

    def __call__(self, feat):
        """Reference layout in, reference layout out: [B,C,h,w] -> [B,1,1,1]."""
        return self.forward_nhwc(feat.permute(0, 2, 3, 1)).clone().view(-1, 1, 1, 1)

    forward = __call__

    def eval(self):
        return self

    def to(self, *a, **k):
        return self
