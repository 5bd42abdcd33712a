"""The sigma-model training step (src/experiments.py:632-700) on libnlc_b200.

One iteration of the reference: perturb the noise and diffuse the batch (:658-669), run the FROZEN UNet's `encode` under
no-grad (:673-682; 99 % of the step's FLOPs), sigma-model forward, loss against the true noise level, backward (:683-691),
AdamW on the master parameters and an EMA copy (:692-694).  Built natively here: the batch preparation (`nlc_train_prepare`,
one pass), the encoder (the sampling path's engine, any of the three network classes), the optimizer + EMA update
(`nlc_adamw_ema_step`, one pass over a flat buffer) and - `NativeSigmaModel` / `train_step_native` - the sigma-model's own
training-mode forward and backward for all three families: DDIM (src/unet_ddim.py:493-529), ADM
(src/unet_adm.py:1029-1083) and EDM (src/edm_networks.py:979-1022).  `SigmaTrainer` / `train_step` keep the first slice: any
`torch.nn.Module` sigma-model under autograd (e.g. the reference's own modules) between the native preparation, encoder and
optimizer.

Data parallelism: the reference wraps the sigma-model in DDP but runs forward and backward under `no_sync()` (:683-687), so
its ranks never average their gradients.  `SigmaTrainer.step` all-reduces the flat gradient buffer over NCCL (one
collective per step, folded mean) before the update, which is what the DDP wrapper was meant to do.
"""
import ctypes as C
import math

import torch

from . import _lib, parallel
from .svd_operators import _stream


def prepare_batch_edm(x0, sigma, noise, extra, eta1, eta2, return_noise=False):
    """The EDM experiment's variant (src/experiments.py:996-1001): new_noise = noise + eta1 (noise + eta2 extra),
    noisy_img = x0 + sigma new_noise, with per-sample sigma [B] (or [B,1,1,1])."""
    return prepare_batch(x0, None, noise, extra, eta1, eta2, None, return_noise=return_noise, sigma=sigma)


def prepare_batch(x0, t, noise, extra, eta1, eta2, alphas_cumprod, return_noise=False, sigma=None):
    """src/experiments.py:661-669 in one kernel.  x0, noise, extra: [B, ...] on the device; t: [B] long; eta1, eta2: [B]
    (or [B,1,1,1]) perturbation factors (`eta1_fn`, `eta2_fn`, :228-231); alphas_cumprod: the scheduler's table.
    Returns (noisy_x, dist_real[, new_noise]): the network input and the regression target ||new_noise|| / sqrt(d)."""
    dev = x0.device
    B = x0.shape[0]
    d = x0[0].numel()
    f = lambda v: v.to(dev, torch.float32).contiguous()
    x0c, nc, ec = f(x0), f(noise), f(extra)
    e1, e2 = f(eta1).reshape(B), f(eta2).reshape(B)
    ab = f(sigma).reshape(B) if sigma is not None else f(alphas_cumprod.to(dev)[t.to(dev).long()]).reshape(B)
    noisy = torch.empty_like(x0c)
    new_noise = torch.empty_like(x0c) if return_noise else None
    dist_real = torch.empty(B, device=dev)
    idx = dev.index if dev.index is not None else torch.cuda.current_device()
    _lib.check(_lib.lib().nlc_train_prepare(_lib.ctx(idx), x0c.data_ptr(), nc.data_ptr(), ec.data_ptr(), e1.data_ptr(),
                                            e2.data_ptr(), ab.data_ptr(), 0 if sigma is None else 1, B, d,
                                            noisy.data_ptr(),
                                            C.c_void_p(new_noise.data_ptr()) if return_noise else None,
                                            dist_real.data_ptr(), _stream()))
    shape = (B,) + (1,) * (x0.dim() - 1)
    out = (noisy, dist_real.view(shape))
    return out + (new_noise,) if return_noise else out


class SigmaTrainer:
    """Optimizer side of `ExperimentDiffusion.set_optimizers` / `train` (:116-170, :692-694) for an fp32 sigma-model: the
    module's parameters are re-pointed into one flat buffer (so the module sees every update), AdamW state and the EMA copy
    are flat buffers of the same length, and `step()` is one all-reduce (if sharded) + one fused kernel."""

    def __init__(self, sigma_model, lr, weight_decay=0.0, betas=(0.9, 0.999), eps=1e-8, ema_rate=0.999):
        self.module = sigma_model
        self.params = [p for p in sigma_model.parameters() if p.requires_grad]
        assert self.params and all(p.is_cuda and p.dtype == torch.float32 for p in self.params), \
            "SigmaTrainer updates fp32 parameters on the GPU (there is no CPU path in this package)"
        dev = self.params[0].device
        sizes = [(p.numel() + 3) // 4 * 4 for p in self.params]  # every tensor starts 16-byte aligned
        n = sum(sizes)
        self.flat = torch.zeros(n, device=dev)
        self.grad = torch.zeros(n, device=dev)
        off = 0
        for p, sz in zip(self.params, sizes):
            view = self.flat[off:off + p.numel()].view_as(p)
            view.copy_(p.data)
            p.data = view
            p.grad = self.grad[off:off + p.numel()].view_as(p)  # autograd accumulates straight into the flat buffer
            off += sz
        self.exp_avg = torch.zeros_like(self.flat)
        self.exp_avg_sq = torch.zeros_like(self.flat)
        self.ema = self.flat.clone()  # copy.deepcopy(master_params), :129
        self.lr, self.weight_decay, self.betas, self.eps, self.ema_rate = lr, weight_decay, betas, eps, ema_rate
        self.steps = 0
        self._ctx = _lib.ctx(dev.index if dev.index is not None else torch.cuda.current_device())

    def zero_grad(self):
        self.grad.zero_()

    def step(self):
        """Average the gradients over the ranks (if any), AdamW, EMA."""
        rank, world = parallel.world()
        if world > 1:
            torch.distributed.all_reduce(self.grad, op=torch.distributed.ReduceOp.SUM)
        self.steps += 1
        _lib.check(_lib.lib().nlc_adamw_ema_step(
            self._ctx, self.flat.data_ptr(), self.grad.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
            self.ema.data_ptr(), self.flat.numel(), self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
            self.steps, self.ema_rate, 1.0 / world, _stream()))

    def ema_state_dict(self):
        """The EMA weights under the module's parameter names (`save_checkpoint`, :244-246)."""
        out, off = {}, 0
        names = [n for n, p in self.module.named_parameters() if p.requires_grad]
        for name, p in zip(names, self.params):
            out[name] = self.ema[off:off + p.numel()].view_as(p).clone()
            off += (p.numel() + 3) // 4 * 4
        return out


def train_step(model, trainer, scheduler, batch_x, t, noise, extra, eta1, eta2, loss_fn, microbatch=None):
    """One iteration of src/experiments.py:654-694.  `model` is an nlc_b200 network (frozen; `encode` runs on the CUDA
    engine), `trainer.module` the sigma-model.  Returns the loss (a device scalar; the reference reads it every step)."""
    noisy_x, dist_real = prepare_batch(batch_x, t, noise, extra, eta1, eta2, scheduler.alphas_cumprod)
    trainer.zero_grad()
    with torch.no_grad():
        B = noisy_x.shape[0]
        mb = microbatch or B
        feat = torch.cat([model.encode(noisy_x[i:i + mb], t[i:i + mb].to(noisy_x.device)).clone()
                          for i in range(0, B, mb)])
    dist_hat = trainer.module(feat) + 1
    loss = loss_fn(dist_real, dist_hat)
    loss.backward()
    trainer.step()
    return loss.detach()


# ======================================================================================================================
# Second slice: the sigma-model's own forward and backward pass, natively (csrc/sigma_train.cu).
class _Flat:
    """Named fp32 tensors as views of one flat buffer, every tensor 16-byte aligned (the layout SigmaTrainer uses)."""

    def __init__(self, shapes, device):
        self.names = list(shapes)
        sizes = [((int(torch.Size(shapes[n]).numel()) + 3) // 4) * 4 for n in self.names]
        self.flat = torch.zeros(sum(sizes), device=device)
        self.views, off = {}, 0
        for n, sz in zip(self.names, sizes):
            self.views[n] = self.flat[off:off + torch.Size(shapes[n]).numel()].view(shapes[n])
            off += sz

    def __getitem__(self, n):
        return self.views[n]


class NativeSigmaModel:
    """The DDIM-family sigma-model (src/unet_ddim.py:493-529: per block PureResnetBlock [-> AttnBlock in block 0] ->
    Downsample; Flatten -> Linear -> BatchNorm1d -> GELU -> Linear) or, with family="adm", the ADM-family one
    (src/unet_adm.py:1029-1083: PureResNetBlock, multi-head AttentionBlock, stride-2 Downsample; the c4 / c5 sigma-model is
    dim 8, 1024 channels, 16 heads), or with family="edm" the EDM one (src/edm_networks.py:979-1022: PureUNetBlock with
    skip_scale, attention in the even blocks, SiLU head; c3: dim 8, 256 channels) with a native training-mode forward AND
    backward pass:
    `loss_and_grad(feat, dist_real)` is `dist_hat = model(feat) + 1; loss = loss_fn(dist_real, dist_hat); loss.backward()`
    of src/experiments.py:688-691 without autograd.  Parameters and gradients live in flat fp32 buffers in the reference's
    parameter layout (`params[name]`, `grads[name]`), which `step()` updates with the fused AdamW + EMA kernel (gradients
    all-reduced over the ranks first).  fp32 throughout; activations NHWC; every contraction is `nlc_sgemm`.
    `dropout` > 0 draws its masks with torch on the device (statistically the reference's nn.Dropout, not its stream)."""

    GROUPS, GN_EPS, BN_EPS, BN_MOMENTUM = 32, 1e-6, 1e-5, 0.1

    def __init__(self, dim=4, channels=64, n_blocks=2, out_dim=1, dropout=0.1, loss="l2", device="cuda", family="ddim",
                 num_heads=1, num_head_channels=-1, use_new_attention_order=False):
        """family "ddim": src/unet_ddim.py:493-529 (single-head AttnBlock with separate q / k / v convolutions, GroupNorm eps
        1e-6, Downsample = pad (0,1,0,1) + stride 2).  family "adm": src/unet_adm.py:1029-1083 (PureResNetBlock with
        in_layers / out_layers, AttentionBlock with one qkv projection and `num_heads` / `num_head_channels` heads in the
        legacy or the new channel order, GroupNorm32 eps 1e-5, Downsample = stride 2 with padding 1).  family "edm":
        src/edm_networks.py:979-1022 (see below; parameters that the reference's forward never applies - `norm1` of a
        PureUNetBlock - keep a zero gradient, where torch's optimizer skips a parameter without gradient: with weight decay the
        fused AdamW shrinks them, which changes nothing the model computes)."""
        if family not in ("ddim", "adm", "edm"):
            raise NotImplementedError("sigma-model family '%s': 'ddim', 'adm' and 'edm' have a native backward" % family)
        self.family = family
        self.heads, self.res_scale, self.head_act = 1, 1.0, 0
        if family == "adm":
            self.GN_EPS = 1e-5
            self.heads = channels // num_head_channels if num_head_channels != -1 else num_heads
            assert channels % self.heads == 0
        elif family == "edm":
            # src/edm_networks.py:979-1022: PureUNetBlock (conv0 feeds conv1 directly; skip_scale sqrt(0.5) after the residual
            # add and after the attention add; one head; GroupNorm with min(32, C / 4) groups, eps 1e-6), attention in the
            # blocks with an even index, the DDIM-style Downsample, and SiLU instead of GELU in the head
            self.GROUPS = min(32, channels // 4)
            self.res_scale, self.head_act = math.sqrt(0.5), 1
        self.new_order = bool(use_new_attention_order)
        self.down_mode = 2 if family == "adm" else 1
        if out_dim != 1:
            raise NotImplementedError("out_dim != 1 is not used by the reference")
        if dim % (2 ** n_blocks) != 0:
            raise NotImplementedError("feature sizes that need the reference's ConstantPad2d (odd sizes) are not built")
        if loss not in ("l2", "l1"):
            raise NotImplementedError("loss_sigma '%s': MSELoss ('l2') and L1Loss ('l1') are built" % loss)
        self.dim, self.C, self.n_blocks, self.p_drop = dim, channels, n_blocks, float(dropout)
        self.loss_kind = 0 if loss == "l2" else 1
        self.device = torch.device(device)
        self._lib, self._bufs = _lib.lib(), {}
        self._ctx = _lib.ctx(self.device.index if self.device.index is not None else torch.cuda.current_device())
        self.training = True
        import os
        self.use_tc = os.environ.get("NLC_TRAIN_TC", "1") != "0"
        self.use_graph = os.environ.get("NLC_GRAPH", "1") != "0"
        self._static = {}

    # ------------------------------------------------------------------ parameters
    def load_state_dict(self, sd, strict=True):
        sd = {k: v.detach() for k, v in sd.items()}
        buffers = ("running_mean", "running_var", "num_batches_tracked")
        shapes = {k: tuple(v.shape) for k, v in sd.items() if not k.endswith(buffers)}
        self.params, self.grads = _Flat(shapes, self.device), _Flat(shapes, self.device)
        for k in shapes:
            self.params[k].copy_(sd[k].to(self.device, torch.float32))
        self.run_mean = sd["fc_layer.2.running_mean"].to(self.device, torch.float32).clone()
        self.run_var = sd["fc_layer.2.running_var"].to(self.device, torch.float32).clone()
        self.num_batches_tracked = int(sd.get("fc_layer.2.num_batches_tracked", 0))
        # module slots: block i owns (pad/identity, resblock, [attn], downsample)
        self.blocks, idx = [], 0
        adm, edm = self.family == "adm", self.family == "edm"
        for i in range(self.n_blocks):
            idx += 1
            r = "down_layer.%d." % idx
            if edm:  # one module: ResBlock part + (even blocks) attention part
                blk = {"norm1": r + "norm0", "conv1": r + "conv0", "conv2": r + "conv1"}
                if i % 2 == 0:
                    blk["attn"] = r
                idx += 1
            else:
                blk = {"norm1": r + ("in_layers.0" if adm else "norm1"), "conv1": r + ("in_layers.2" if adm else "conv1"),
                       "norm2": r + ("out_layers.0" if adm else "norm2"), "conv2": r + ("out_layers.3" if adm else "conv2")}
                idx += 1
                if i == 0:
                    blk["attn"] = "down_layer.%d." % idx
                    idx += 1
            blk["down"] = "down_layer.%d." % idx + ("op" if adm else "conv")
            idx += 1
            self.blocks.append(blk)
        need = [self.blocks[0]["conv1"] + ".weight", self.blocks[0]["attn"] + ("q.weight" if self.family == "ddim" else "qkv.weight"),
                "fc_layer.1.weight", "final_mlp.weight"]
        missing = [k for k in need if k not in shapes]
        if missing:
            raise KeyError("sigma-model state_dict lacks %s" % missing)
        self.fc_dim = shapes["fc_layer.1.weight"][0]
        self.exp_avg = torch.zeros_like(self.params.flat)
        self.exp_avg_sq = torch.zeros_like(self.params.flat)
        self.ema = self.params.flat.clone()
        self.steps = 0
        self._static = {}  # (captured passes point at the previous parameter buffers)
        return self

    def state_dict(self):
        sd = {k: self.params[k].clone() for k in self.params.names}
        sd["fc_layer.2.running_mean"], sd["fc_layer.2.running_var"] = self.run_mean.clone(), self.run_var.clone()
        sd["fc_layer.2.num_batches_tracked"] = torch.tensor(self.num_batches_tracked)
        return sd

    def ema_state_dict(self):
        off, out = 0, {}
        for k in self.params.names:
            n = self.params[k].numel()
            out[k] = self.ema[off:off + n].view_as(self.params[k]).clone()
            off += (n + 3) // 4 * 4
        return out

    # ------------------------------------------------------------------ plumbing
    def _buf(self, tag, *shape):
        key = (tag,) + tuple(shape)
        t = self._bufs.get(key)
        if t is None:
            t = torch.empty(*shape, device=self.device)
            self._bufs[key] = t
        return t

    def _mm(self, batch, M, N, K, A, sa, B, sb, out, add=None):
        """out[b] = A[b] B[b] (+ add[b]); sa = (batch, row, k) strides of A, sb = (batch, k, column) strides of B."""
        _lib.check(self._lib.nlc_sgemm(self._ctx, batch, M, N, K, A.data_ptr(), sa[0], sa[1], sa[2], B.data_ptr(), sb[0], sb[1],
                                       sb[2], out.data_ptr(), C.c_void_p(add.data_ptr()) if add is not None else None, _stream()))

    # The large contractions (the 3x3 convolutions' forward / data-gradient / weight-gradient GEMMs: 98 % of the sigma-model's
    # FLOPs) go to the tensor cores in the fp32-accurate mode of nlc_conv_tc (NLC_F32X3: three tf32 split products, per-chunk
    # accumulators): out[M, N] = A[M, K] Bt[N, K]^T with both operands K-major, M % 128 == 0, N % 64 == 0, K % 32 == 0.  The
    # CUDA-core nlc_sgemm serves the rest (attention products, the head, shapes that do not tile) - and everything when
    # NLC_TRAIN_TC=0.
    TC_MIN_FLOP = 1 << 28

    def _tc_ok(self, M, N, K):
        return self.use_tc and M % 128 == 0 and N % 64 == 0 and K % 32 == 0 and 2 * M * N * K >= self.TC_MIN_FLOP

    def _mm_tc(self, M, N, K, A, Bt, out, bias=None):
        """out[M, N] = A[M, K] Bt[N, K]^T (+ bias[N]) on tcgen05, fp32-accurate (A, Bt, out contiguous fp32)."""
        from . import ops
        from ._lib import NLC_F32X3
        ops.conv_tc([ops.Act(A.view(1, M // 128, 128, K))], [(0, 0, 0, 0, K)], Bt, N, 1, M // 128, 128, NLC_F32X3, bias=bias,
                    out_f32=ops.Act(out.view(1, M // 128, 128, N)))

    def _transpose(self, x, rows, cols, tag):
        """[rows, cols] -> [cols, rows] (contiguous fp32 scratch)."""
        y = self._buf(tag, cols, rows)
        _lib.check(self._lib.nlc_permute_nhwc(self._ctx, x.data_ptr(), 1, rows, cols, 1, y.data_ptr(), _stream()))
        return y

    def _linear(self, x, rows, cin, w, b, out):
        """out[rows, cout] = x[rows, cin] w[cout, cin]^T + b"""
        cout = w.shape[0]
        if self._tc_ok(rows, cout, cin):
            return self._mm_tc(rows, cout, cin, x, w, out, bias=b)
        self._mm(1, rows, cout, cin, x, (0, cin, 1), w, (0, 1, cin), out)
        _lib.check(self._lib.nlc_bias_add(self._ctx, out.data_ptr(), b.data_ptr(), rows, cout, _stream()))

    def _linear_bwd(self, x, dy, rows, cin, w, dw, db, dx, add=None):
        """dw[cout, cin] = dy^T x; db = column sums of dy; dx[rows, cin] = dy w (+ add)"""
        cout = w.shape[0]
        _lib.check(self._lib.nlc_colsum(self._ctx, dy.data_ptr(), rows, cout, db.data_ptr(), _stream()))
        if self._tc_ok(cout, cin, rows):
            dyt = self._transpose(dy, rows, cout, "tc.dyt")     # [cout, rows]
            xt = self._transpose(x, rows, cin, "tc.xt")         # [cin, rows]
            self._mm_tc(cout, cin, rows, dyt, xt, dw)
        else:
            self._mm(1, cout, cin, rows, dy, (0, 1, cout), x, (0, cin, 1), dw)
        if dx is not None:
            if self._tc_ok(rows, cin, cout) and add is None:
                wt = self._transpose(w, cout, cin, "tc.wt")     # [cin, cout]
                self._mm_tc(rows, cin, cout, dy, wt, dx)
            else:
                self._mm(1, rows, cin, cout, dy, (0, cout, 1), w, (0, cin, 1), dx, add=add)

    def _gn(self, x, B, HW, name, act, y, stats):
        _lib.check(self._lib.nlc_gn_train_fwd(self._ctx, x.data_ptr(), B, HW, self.C, self.GROUPS, self.GN_EPS,
                                              self.params[name + ".weight"].data_ptr(), self.params[name + ".bias"].data_ptr(), act,
                                              y.data_ptr(), stats.data_ptr(), _stream()))

    def _gn_bwd(self, x, dy, B, HW, name, act, stats, dx, accumulate):
        _lib.check(self._lib.nlc_gn_train_bwd(self._ctx, x.data_ptr(), dy.data_ptr(), B, HW, self.C, self.GROUPS,
                                              self.params[name + ".weight"].data_ptr(), self.params[name + ".bias"].data_ptr(), act,
                                              stats.data_ptr(), dx.data_ptr(), 1 if accumulate else 0,
                                              self.grads[name + ".weight"].data_ptr(), self.grads[name + ".bias"].data_ptr(),
                                              _stream()))

    def _conv3(self, x, B, H, W, name, down, tag):
        """3x3 conv through its patch matrix; returns (y [B*Ho*Wo, C], patches)."""
        Cc = self.C
        Ho, Wo = (H // 2, W // 2) if down else (H, W)
        M = B * Ho * Wo
        P = self._buf(tag + ".P", M, 9 * Cc)
        _lib.check(self._lib.nlc_unfold3x3(self._ctx, x.data_ptr(), B, H, W, Cc, down, P.data_ptr(), _stream()))
        y = self._buf(tag + ".y", M, Cc)
        self._linear(P, M, 9 * Cc, self.params[name + ".weight"].view(Cc, 9 * Cc), self.params[name + ".bias"], y)
        return y, P

    def _conv3_bwd(self, dy, P, B, H, W, name, down, dx, beta, tag):
        Cc = self.C
        Ho, Wo = (H // 2, W // 2) if down else (H, W)
        M = B * Ho * Wo
        dP = self._buf(tag + ".dP", M, 9 * Cc)
        self._linear_bwd(P, dy, M, 9 * Cc, self.params[name + ".weight"].view(Cc, 9 * Cc),
                         self.grads[name + ".weight"].view(Cc, 9 * Cc), self.grads[name + ".bias"], dP)
        _lib.check(self._lib.nlc_fold3x3(self._ctx, dP.data_ptr(), B, H, W, Cc, down, dx.data_ptr(), beta, _stream()))

    def _head_view(self, qkv, B, HW, nh, ch, which):
        """[B, T, heads, ch] view of q (0) / k (1) / v (2) inside a [B*T, 3C] projection: legacy order = per head
        [q | k | v] (src/unet_adm.py:340-345), new order = [q | k | v] x [head] (:373-377)."""
        if self.family == "edm":  # reshape(B * heads, ch, 3, T).unbind(2): channel = (head * ch + c) * 3 + which
            return qkv.view(B, HW, nh, ch, 3)[..., which]
        if self.new_order:
            return qkv.view(B, HW, 3, nh, ch)[:, :, which]
        return qkv.view(B, HW, nh, 3, ch)[:, :, :, which]

    def _axpby(self, a, x, b, y, out):
        _lib.check(self._lib.nlc_axpby(self._ctx, a, x.data_ptr(), b, C.c_void_p(y.data_ptr()) if y is not None else None,
                                       out.data_ptr(), out.numel(), _stream()))

    # ------------------------------------------------------------------ forward + backward
    def loss_and_grad(self, feat, dist_real, nhwc=False, weight=None):
        """feat: encoder feature [B, C, dim, dim] (the reference's layout) or NHWC [B, dim, dim, C] with nhwc=True;
        dist_real: [B] (or [B,1,1,1]) targets.  Training-mode forward, loss, backward: fills `grads`, updates the BatchNorm
        running statistics, returns (loss [1], dist_hat [B]) on the device (plan buffers: valid until the next call).
        The ~130 launches of the pass are captured in a CUDA graph after the first call at a batch size (NLC_GRAPH=0:
        always eager) - at 4x4 .. 8x8 pixels the pass is launch-bound otherwise."""
        B = feat.shape[0]
        key = (B, bool(nhwc), weight is not None)
        st = self._static.get(key)
        if st is None:
            st = dict(feat=torch.empty(tuple(feat.shape), device=self.device), target=torch.empty(B, device=self.device),
                      weight=torch.empty(B, device=self.device) if weight is not None else None, graph=None)
            self._static[key] = st
        st["feat"].copy_(feat)
        st["target"].copy_(dist_real.reshape(B))
        if weight is not None:  # per-sample loss weights (normalised by their sum inside the loss kernel)
            st["weight"].copy_(weight.reshape(B))
        if st["graph"] is not None:
            st["graph"].replay()
        else:
            out = self._loss_and_grad(st["feat"], st["target"], nhwc, st["weight"])
            st["out"] = out
            if self.use_graph and self.device.type == "cuda":
                torch.cuda.synchronize()
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    st["out"] = self._loss_and_grad(st["feat"], st["target"], nhwc, st["weight"])
                st["graph"] = g
        self.num_batches_tracked += 1
        return st["out"]

    def _loss_and_grad(self, feat, target, nhwc, weight=None):
        L, ctx, Cc = self._lib, self._ctx, self.C
        B, dim = feat.shape[0], self.dim
        self.grads.flat.zero_()
        x = self._buf("in", B * dim * dim, Cc)
        if nhwc:
            x.copy_(feat.reshape(B * dim * dim, Cc))
        else:
            _lib.check(L.nlc_permute_nhwc(ctx, feat.data_ptr(), B, dim * dim, Cc, 0, x.data_ptr(), _stream()))
        saved, H = [], dim
        for bi, blk in enumerate(self.blocks):
            M, HW = B * H * H, H * H
            t = "b%d" % bi
            a1, st1 = self._buf(t + ".a1", M, Cc), self._buf(t + ".st1", B * self.GROUPS * 2)
            self._gn(x, B, HW, blk["norm1"], 1, a1, st1)
            c1, P1 = self._conv3(a1, B, H, H, blk["conv1"], 0, t + ".c1")
            rs = self.res_scale
            if self.family == "edm":  # PureUNetBlock: conv1(dropout(conv0(silu(norm0 x)))), no second normalisation
                a2, st2 = c1, None
            else:
                a2, st2 = self._buf(t + ".a2", M, Cc), self._buf(t + ".st2", B * self.GROUPS * 2)
                self._gn(c1, B, HW, blk["norm2"], 1, a2, st2)
            mask = None
            if self.training and self.p_drop > 0:
                mask = (torch.rand(M, Cc, device=self.device) >= self.p_drop).float() / (1.0 - self.p_drop)
                a2.mul_(mask)
            c2, P2 = self._conv3(a2, B, H, H, blk["conv2"], 0, t + ".c2")
            y = self._buf(t + ".y", M, Cc)
            self._axpby(rs, x, rs, c2, y)  # (x + h) * skip_scale; skip_scale = 1 outside the EDM family
            rec = dict(x=x, st1=st1, P1=P1, c1=c1, st2=st2, P2=P2, mask=mask, H=H)
            z = y
            if "attn" in blk and self.family != "ddim":
                # AttentionBlock (src/unet_adm.py:299-305) / the attention half of a PureUNetBlock (src/edm_networks.py:
                # 948-954): one qkv projection, `heads` heads of ch channels; the heads are regrouped into [B * heads, T, ch]
                # operands (torch copies: data movement only) around the batched GEMMs
                a, nh = blk["attn"], self.heads
                ch = Cc // nh
                anorm, aproj = (a + "norm2", a + "proj") if self.family == "edm" else (a + "norm", a + "proj_out")
                n, stn = self._buf(t + ".n", M, Cc), self._buf(t + ".stn", B * self.GROUPS * 2)
                self._gn(y, B, HW, anorm, 0, n, stn)
                qkv = self._buf(t + ".qkv", M, 3 * Cc)
                self._linear(n, M, Cc, self.params[a + "qkv.weight"].view(3 * Cc, Cc), self.params[a + "qkv.bias"], qkv)
                q, k, v = (self._buf(t + "." + nm + "h", B * nh, HW, ch) for nm in "qkv")
                for i, dst in enumerate((q, k, v)):
                    dst.view(B, nh, HW, ch).copy_(self._head_view(qkv, B, HW, nh, ch, i).permute(0, 2, 1, 3))
                S, Pm = self._buf(t + ".S", B * nh, HW, HW), self._buf(t + ".Pm", B * nh, HW, HW)
                self._mm(B * nh, HW, HW, ch, q, (HW * ch, ch, 1), k, (HW * ch, 1, ch), S)
                scale = float(1.0 / math.sqrt(ch))  # (q ch^-1/4) . (k ch^-1/4), src/unet_adm.py:346-349
                _lib.check(L.nlc_softmax_rows(ctx, S.data_ptr(), None, B * nh * HW, HW, scale, Pm.data_ptr(), _stream()))
                Oh = self._buf(t + ".Oh", B * nh, HW, ch)
                self._mm(B * nh, HW, ch, HW, Pm, (HW * HW, HW, 1), v, (HW * ch, ch, 1), Oh)
                O = self._buf(t + ".O", M, Cc)
                O.view(B, HW, nh, ch).copy_(Oh.view(B, nh, HW, ch).permute(0, 2, 1, 3))
                pr = self._buf(t + ".pr", M, Cc)
                self._linear(O, M, Cc, self.params[aproj + ".weight"].view(Cc, Cc), self.params[aproj + ".bias"], pr)
                z = self._buf(t + ".z", M, Cc)
                self._axpby(rs, y, rs, pr, z)
                rec.update(y=y, n=n, stn=stn, q=q, k=k, v=v, Pm=Pm, O=O, scale=scale)
            elif "attn" in blk:
                a = blk["attn"]
                n, stn = self._buf(t + ".n", M, Cc), self._buf(t + ".stn", B * self.GROUPS * 2)
                self._gn(y, B, HW, a + "norm", 0, n, stn)
                q, k, v = (self._buf(t + "." + nm, M, Cc) for nm in "qkv")
                for nm, dst in (("q", q), ("k", k), ("v", v)):
                    self._linear(n, M, Cc, self.params[a + nm + ".weight"].view(Cc, Cc), self.params[a + nm + ".bias"], dst)
                S, Pm = self._buf(t + ".S", B, HW, HW), self._buf(t + ".Pm", B, HW, HW)
                self._mm(B, HW, HW, Cc, q, (HW * Cc, Cc, 1), k, (HW * Cc, 1, Cc), S)
                scale = float(int(Cc) ** (-0.5))
                _lib.check(L.nlc_softmax_rows(ctx, S.data_ptr(), None, B * HW, HW, scale, Pm.data_ptr(), _stream()))
                O = self._buf(t + ".O", M, Cc)
                self._mm(B, HW, Cc, HW, Pm, (HW * HW, HW, 1), v, (HW * Cc, Cc, 1), O)
                pr = self._buf(t + ".pr", M, Cc)
                self._linear(O, M, Cc, self.params[a + "proj_out.weight"].view(Cc, Cc), self.params[a + "proj_out.bias"], pr)
                z = self._buf(t + ".z", M, Cc)
                self._axpby(1.0, y, 1.0, pr, z)
                rec.update(y=y, n=n, stn=stn, q=q, k=k, v=v, Pm=Pm, O=O, scale=scale)
            d, Pd = self._conv3(z, B, H, H, blk["down"], self.down_mode, t + ".d")
            rec.update(Pd=Pd)
            saved.append(rec)
            x, H = d, H // 2
        HWf, hidden = H * H, Cc * H * H
        hflat = self._buf("hflat", B, hidden)
        _lib.check(L.nlc_permute_nhwc(ctx, x.data_ptr(), B, HWf, Cc, 1, hflat.data_ptr(), _stream()))
        F1 = self.fc_dim
        f1, g, stb = self._buf("f1", B, F1), self._buf("g", B, F1), self._buf("stb", F1 * 2)
        self._linear(hflat, B, hidden, self.params["fc_layer.1.weight"], self.params["fc_layer.1.bias"], f1)
        _lib.check(L.nlc_bn1d_act_train(ctx, f1.data_ptr(), None, B, F1, self.BN_EPS, self.BN_MOMENTUM, self.head_act,
                                        self.params["fc_layer.2.weight"].data_ptr(), self.params["fc_layer.2.bias"].data_ptr(),
                                        self.run_mean.data_ptr(), self.run_var.data_ptr(), stb.data_ptr(), g.data_ptr(), None,
                                        None, _stream()))
        r = self._buf("r", B, 1)
        self._linear(g, B, F1, self.params["final_mlp.weight"], self.params["final_mlp.bias"], r)
        dist_hat, loss, dr = self._buf("dist_hat", B), self._buf("loss", 1), self._buf("dr", B, 1)
        if weight is None:
            _lib.check(L.nlc_head_loss(ctx, r.data_ptr(), target.data_ptr(), B, self.loss_kind, dist_hat.data_ptr(),
                                       loss.data_ptr(), dr.data_ptr(), _stream()))
        else:
            _lib.check(L.nlc_head_loss_weighted(ctx, r.data_ptr(), target.data_ptr(), weight.data_ptr(), B, self.loss_kind,
                                                dist_hat.data_ptr(), loss.data_ptr(), dr.data_ptr(), _stream()))
        # ---------------------------------------------------------------- backward
        dg = self._buf("dg", B, F1)
        self._linear_bwd(g, dr, B, F1, self.params["final_mlp.weight"], self.grads["final_mlp.weight"],
                         self.grads["final_mlp.bias"], dg)
        df1 = self._buf("df1", B, F1)
        _lib.check(L.nlc_bn1d_act_train(ctx, f1.data_ptr(), dg.data_ptr(), B, F1, self.BN_EPS, self.BN_MOMENTUM, self.head_act,
                                        self.params["fc_layer.2.weight"].data_ptr(), self.params["fc_layer.2.bias"].data_ptr(),
                                        None, None, stb.data_ptr(), df1.data_ptr(), self.grads["fc_layer.2.weight"].data_ptr(),
                                        self.grads["fc_layer.2.bias"].data_ptr(), _stream()))
        dh = self._buf("dh", B, hidden)
        self._linear_bwd(hflat, df1, B, hidden, self.params["fc_layer.1.weight"], self.grads["fc_layer.1.weight"],
                         self.grads["fc_layer.1.bias"], dh)
        dx = self._buf("dlast", B * HWf, Cc)
        _lib.check(L.nlc_permute_nhwc(ctx, dh.data_ptr(), B, HWf, Cc, 0, dx.data_ptr(), _stream()))
        for bi in reversed(range(len(self.blocks))):
            blk, rec, t = self.blocks[bi], saved[bi], "b%d" % bi
            H = rec["H"]
            M, HW = B * H * H, H * H
            dz = self._buf(t + ".dz", M, Cc)
            self._conv3_bwd(dx, rec["Pd"], B, H, H, blk["down"], self.down_mode, dz, 0.0, t + ".d")
            dy = dz
            rs = self.res_scale
            if "attn" in blk and self.family != "ddim":
                a, Pm, scale, nh = blk["attn"], rec["Pm"], rec["scale"], self.heads
                ch = Cc // nh
                anorm, aproj = (a + "norm2", a + "proj") if self.family == "edm" else (a + "norm", a + "proj_out")
                if rs != 1.0:  # z = (y + proj) * skip_scale
                    dzs = self._buf(t + ".dzs", M, Cc)
                    self._axpby(rs, dz, 0.0, None, dzs)
                    dz = dzs
                dO = self._buf(t + ".dO", M, Cc)
                self._linear_bwd(rec["O"], dz, M, Cc, self.params[aproj + ".weight"].view(Cc, Cc),
                                 self.grads[aproj + ".weight"].view(Cc, Cc), self.grads[aproj + ".bias"], dO)
                dOh = self._buf(t + ".dOh", B * nh, HW, ch)
                dOh.view(B, nh, HW, ch).copy_(dO.view(B, HW, nh, ch).permute(0, 2, 1, 3))
                dPm, dS = self._buf(t + ".dPm", B * nh, HW, HW), self._buf(t + ".dS", B * nh, HW, HW)
                self._mm(B * nh, HW, HW, ch, dOh, (HW * ch, ch, 1), rec["v"], (HW * ch, 1, ch), dPm)
                dv, dq, dk = (self._buf(t + ".d" + nm + "h", B * nh, HW, ch) for nm in "vqk")
                self._mm(B * nh, HW, ch, HW, Pm, (HW * HW, 1, HW), dOh, (HW * ch, ch, 1), dv)
                _lib.check(L.nlc_softmax_rows(ctx, Pm.data_ptr(), dPm.data_ptr(), B * nh * HW, HW, scale, dS.data_ptr(), _stream()))
                self._mm(B * nh, HW, ch, HW, dS, (HW * HW, HW, 1), rec["k"], (HW * ch, ch, 1), dq)
                self._mm(B * nh, HW, ch, HW, dS, (HW * HW, 1, HW), rec["q"], (HW * ch, ch, 1), dk)
                dqkv = self._buf(t + ".dqkv", M, 3 * Cc)
                for i, src in enumerate((dq, dk, dv)):
                    self._head_view(dqkv, B, HW, nh, ch, i).copy_(src.view(B, nh, HW, ch).permute(0, 2, 1, 3))
                dn = self._buf(t + ".dn", M, Cc)
                self._linear_bwd(rec["n"], dqkv, M, Cc, self.params[a + "qkv.weight"].view(3 * Cc, Cc),
                                 self.grads[a + "qkv.weight"].view(3 * Cc, Cc), self.grads[a + "qkv.bias"], dn)
                dy = self._buf(t + ".dy", M, Cc)
                dy.copy_(dz)
                self._gn_bwd(rec["y"], dn, B, HW, anorm, 0, rec["stn"], dy, True)
            elif "attn" in blk:
                a, Pm, scale = blk["attn"], rec["Pm"], rec["scale"]
                dO = self._buf(t + ".dO", M, Cc)
                self._linear_bwd(rec["O"], dz, M, Cc, self.params[a + "proj_out.weight"].view(Cc, Cc),
                                 self.grads[a + "proj_out.weight"].view(Cc, Cc), self.grads[a + "proj_out.bias"], dO)
                dPm, dS = self._buf(t + ".dPm", B, HW, HW), self._buf(t + ".dS", B, HW, HW)
                self._mm(B, HW, HW, Cc, dO, (HW * Cc, Cc, 1), rec["v"], (HW * Cc, 1, Cc), dPm)
                dv, dq, dk = (self._buf(t + ".d" + nm, M, Cc) for nm in "vqk")
                self._mm(B, HW, Cc, HW, Pm, (HW * HW, 1, HW), dO, (HW * Cc, Cc, 1), dv)
                _lib.check(L.nlc_softmax_rows(ctx, Pm.data_ptr(), dPm.data_ptr(), B * HW, HW, scale, dS.data_ptr(), _stream()))
                self._mm(B, HW, Cc, HW, dS, (HW * HW, HW, 1), rec["k"], (HW * Cc, Cc, 1), dq)
                self._mm(B, HW, Cc, HW, dS, (HW * HW, 1, HW), rec["q"], (HW * Cc, Cc, 1), dk)
                dn = self._buf(t + ".dn", M, Cc)
                first = True
                for nm, dsrc in (("q", dq), ("k", dk), ("v", dv)):
                    self._linear_bwd(rec["n"], dsrc, M, Cc, self.params[a + nm + ".weight"].view(Cc, Cc),
                                     self.grads[a + nm + ".weight"].view(Cc, Cc), self.grads[a + nm + ".bias"], dn,
                                     add=None if first else dn)
                    first = False
                dy = self._buf(t + ".dy", M, Cc)
                dy.copy_(dz)
                self._gn_bwd(rec["y"], dn, B, HW, a + "norm", 0, rec["stn"], dy, True)
            # ResBlock: y = x + conv2(drop(swish(gn2(conv1(swish(gn1(x)))))))
            if rs != 1.0:  # y = (x + conv2(...)) * skip_scale: both branches see skip_scale * dy
                dys = self._buf(t + ".dys", M, Cc)
                self._axpby(rs, dy, 0.0, None, dys)
                dy = dys
            da2 = self._buf(t + ".da2", M, Cc)
            self._conv3_bwd(dy, rec["P2"], B, H, H, blk["conv2"], 0, da2, 0.0, t + ".c2")
            if rec["mask"] is not None:
                da2.mul_(rec["mask"])
            if self.family == "edm":
                dc1 = da2
            else:
                dc1 = self._buf(t + ".dc1", M, Cc)
                self._gn_bwd(rec["c1"], da2, B, HW, blk["norm2"], 1, rec["st2"], dc1, False)
            da1 = self._buf(t + ".da1", M, Cc)
            self._conv3_bwd(dc1, rec["P1"], B, H, H, blk["conv1"], 0, da1, 0.0, t + ".c1")
            dxin = self._buf(t + ".dx", M, Cc)
            dxin.copy_(dy)
            self._gn_bwd(rec["x"], da1, B, HW, blk["norm1"], 1, rec["st1"], dxin, True)
            dx = dxin
        return loss, dist_hat

    # ------------------------------------------------------------------ optimizer
    def step(self, lr, weight_decay=0.0, betas=(0.9, 0.999), eps=1e-8, ema_rate=0.999):
        """All-reduce the gradients over the ranks (mean), AdamW + EMA in one fused pass (nlc_adamw_ema_step)."""
        rank, world = parallel.world()
        if world > 1:
            torch.distributed.all_reduce(self.grads.flat, op=torch.distributed.ReduceOp.SUM)
        self.steps += 1
        _lib.check(self._lib.nlc_adamw_ema_step(
            self._ctx, self.params.flat.data_ptr(), self.grads.flat.data_ptr(), self.exp_avg.data_ptr(),
            self.exp_avg_sq.data_ptr(), self.ema.data_ptr(), self.params.flat.numel(), lr, betas[0], betas[1], eps, weight_decay,
            self.steps, ema_rate, 1.0 / world, _stream()))


def train_step_native(model, sigma, scheduler, batch_x, t, noise, extra, eta1, eta2, lr, weight_decay=0.0, microbatch=None,
                      **adam):
    """One iteration of src/experiments.py:654-694 with nothing left to autograd: batch preparation kernel, frozen-UNet
    `encode_scaled` on the tensor-core engine, `NativeSigmaModel.loss_and_grad`, fused AdamW + EMA.  Returns the loss."""
    noisy_x, dist_real = prepare_batch(batch_x, t, noise, extra, eta1, eta2, scheduler.alphas_cumprod)
    B = noisy_x.shape[0]
    mb = microbatch or B
    feats = [model.encode_scaled(noisy_x[i:i + mb], t[i:i + mb].to(noisy_x.device)).clone() for i in range(0, B, mb)]
    loss, _ = sigma.loss_and_grad(torch.cat(feats), dist_real, nhwc=True)
    sigma.step(lr, weight_decay=weight_decay, **adam)
    return loss


def train_step_native_edm(model, sigma_model, batch_x, sigma, noise, extra, eta1, eta2, lr, sigma_data=0.5, weight_decay=0.0,
                          microbatch=None, loss_weighted=False, **adam):
    """One iteration of the EDM experiment's training loop (src/experiments.py:990-1024; `loss_weighted` = its per-sample
    weights (sigma^2 + sigma_data^2) / (sigma sigma_data)^2 normalised by their sum, :994,1019-1021) with nothing left to
    autograd: `prepare_batch_edm` (noisy_img = x + sigma new_noise, dist_real), the frozen SongUNet's `encode` with the EDM
    preconditioning of `encode_edm` (c_in = 1 / sqrt(sigma_data^2 + sigma^2), c_noise = log(sigma) / 4, :777-786) on the
    tensor-core engine, `NativeSigmaModel(family="edm").loss_and_grad`, fused AdamW + EMA.  sigma: [B] (or [B,1,1,1]) noise
    levels, drawn by the caller as the reference does (:990-993).  Returns the loss (device scalar)."""
    noisy_x, dist_real = prepare_batch_edm(batch_x, sigma, noise, extra, eta1, eta2)
    B = noisy_x.shape[0]
    sg = sigma.to(noisy_x.device, torch.float32).reshape(B)
    c_in = 1.0 / (sigma_data ** 2 + sg ** 2).sqrt()
    c_noise = sg.log() / 4
    mb = microbatch or B
    feats = [model.encode_scaled(noisy_x[i:i + mb], c_noise[i:i + mb].contiguous(), c_in[i:i + mb].contiguous()).clone()
             for i in range(0, B, mb)]
    weight = (sg ** 2 + sigma_data ** 2) / (sg * sigma_data) ** 2 if loss_weighted else None
    loss, _ = sigma_model.loss_and_grad(torch.cat(feats), dist_real, nhwc=True, weight=weight)
    sigma_model.step(lr, weight_decay=weight_decay, **adam)
    return loss

