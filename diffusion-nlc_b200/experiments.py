"""Host-side mirror of the reference's sampler orchestration (src/experiments.py: ExperimentDiffusion,
ImageExperiment) for the sampling path: `denoise_loop` and `get_denoise_vector` keep the reference's names,
arguments and return values, and every per-step tensor operation is a libnlc_b200 kernel.

Per NLC step (src/experiments.py:400-460):
    row_norm -> refine_sigma (clamp + searchsorted) -> UNet.encode (input scale folded into conv_in)
    -> sigma-model -> sigma_correct (sigma_hat, sigma_prev_hat, t_hat) -> UNet.forward -> normalize_rows
then pred_xstart(+clip) -> [constraint projection] -> pred_xprev (src/experiments.py:346-390).

Differences from the reference, all host-side and documented in DESIGN.md:
  * `chunk_size` micro-batching (a memory-saving device of the reference) is accepted and ignored: samples are
    independent, so results are identical;
  * the NaN early-exit (`torch.isnan(xt).any()`, one host sync per step) reads a device flag every
    `nan_check_every` steps instead;
  * `return_log=True` returns the same lists but costs the same device->host copies as in the reference.
"""
import math
import os

import numpy as np
import torch

from . import ops
from .schedulers import CLIP_CLAMP, CLIP_NONE, _as_f32

CLIP_DYNAMIC = 2


class _Work:
    """Per-batch-size scratch vectors of the step (all fp32, device)."""

    def __init__(self, B, shape, device):
        f = lambda *s: torch.empty(*s, device=device, dtype=torch.float32)
        self.norms, self.sigma, self.t, self.scale = f(B), f(B), f(B), f(B)
        self.sigma_hat, self.sigma_prev_hat, self.t_hat, self.scale_hat = f(B), f(B), f(B), f(B)
        self.sigma_prev = f(B)
        self.eps = f(B, *shape)
        self.x0 = f(B, *shape)
        self.xa, self.xb = f(B, *shape), f(B, *shape)
        self.nan_flag = torch.zeros(1, device=device, dtype=torch.int32)
        self._shape, self._device = (B,) + tuple(shape), device
        self._graph_bufs = None

    def graph_bufs(self):
        """Fixed-address inputs of a captured step: (sigma_t, sigma_prev) of the step, its noise draw, and the device-side
        best-x0 bookkeeping (value, flag, image)."""
        if self._graph_bufs is None:
            import types
            f = lambda *s: torch.empty(*s, device=self._device, dtype=torch.float32)
            self._graph_bufs = types.SimpleNamespace(
                step_sig=f(2), noise=f(*self._shape), best_val=f(1), best_x0=f(*self._shape), loss_sum=f(1),
                flag=torch.zeros(1, device=self._device, dtype=torch.int32))
        return self._graph_bufs


class ExperimentDiffusion:
    """src/experiments.py:87-551 (sampling methods)."""

    def __init__(self, model, scheduler, batch_size, data_shape, save_folder, seed=0, device="cuda", dist_train=0,
                 time_shift=0):
        self.model = model
        self.scheduler = scheduler
        self.device = torch.device(device)
        self.seed = seed
        self.batch_size = batch_size
        self.data_shape = tuple(data_shape)
        self.shape = (batch_size,) + self.data_shape
        self.dim = int(np.prod(data_shape))
        self.dim_coord = len(data_shape)
        self.save_folder = save_folder
        self.dist_train = dist_train
        self.time_shift = time_shift
        self.clip_mode = CLIP_NONE
        self.learn_epsvar = False
        self.sigma_model = None
        self.norm_min, self.norm_max = 0.0, 1.0
        self.nan_check_every = 16
        # replay the timestep from a CUDA graph in denoise_loop (NLC_GRAPH=0 disables; see denoise_loop)
        self.cuda_graph = os.environ.get("NLC_GRAPH", "1") != "0"
        self._graph_cache = {}  # captured timesteps, kept across denoise_loop calls (see _graph_signature)
        # instrumentation for parity studies: `time_source(step) -> (t, t_hat)` ([B] float tensors or None) replaces the
        # step's two discrete time lookups t = searchsorted(sigma) (src/experiments.py:410,427) by given values, so a
        # free-running trajectory can be compared with the reference's without its time-bucket decisions diverging
        self.time_source = None
        self._step = 0
        self.gen = self.new_gen()
        self._work = {}

    # ---------------------------------------------------------------- configuration (same names as the reference)
    def set_model(self, model=None, sigma_model=None, learn_epsvar=True):
        if model is not None:
            self.model = model
            self.learn_epsvar = learn_epsvar
        else:
            self.learn_epsvar = False
        if sigma_model is not None:
            self.sigma_model = sigma_model

    def set_norm_maxmin(self, norm_min=None, norm_max=None):
        # src/experiments.py:176-184
        self.norm_min = norm_min / math.sqrt(self.dim) if norm_min is not None else 0.0
        self.norm_max = norm_max / math.sqrt(self.dim) if norm_max is not None else 1.0

    def set_clip_fn(self, clip_fn="none"):
        # src/experiments.py:186-207
        if clip_fn == "clamp":
            self.clip_mode = CLIP_CLAMP
        elif clip_fn == "dynamic":
            # partial(_threshold_sample, dynamic_thresholding_ratio=0.99, sample_max_value=100), :204
            self.clip_mode = CLIP_DYNAMIC
            self.dynamic_ratio, self.dynamic_max = 0.99, 100.0
        else:
            self.clip_mode = CLIP_NONE

    def new_gen(self, seed=None):
        return torch.manual_seed(self.seed if seed is None else seed)

    def get_noise(self, shape=None, gen=None, norm_noise=False):
        # src/experiments.py:263-271: CPU generator, then one host->device copy
        shape = self.shape if shape is None else shape
        gen = self.gen if gen is None else gen
        noise = torch.randn(shape, generator=gen).to(self.device)
        if norm_noise:
            ops.normalize_rows_(noise)
        return noise

    def get_noise_xt(self, shape=None, gen=None, norm_noise=False, t=None, sigma=None):
        # x_T = z / sqrt(alpha_bar) with alpha_bar = 1/(sigma^2+1)  (src/experiments.py:284-293,322-325)
        zt = self.get_noise(shape=shape, gen=gen, norm_noise=norm_noise)
        alpha_bar = 1 / (sigma.to(self.device) ** 2 + 1)
        return zt / alpha_bar.sqrt(), zt

    def convert_coordinate(self, xt, sigma):
        return xt * (1 / (sigma ** 2 + 1)).sqrt()

    def _w(self, B):
        w = self._work.get(B)
        if w is None:
            w = _Work(B, self.data_shape, self.device)
            self._work[B] = w
        return w

    # ---------------------------------------------------------------- D1: the fused step front-end
    @torch.no_grad()
    def get_denoise_vector(self, xt, t, sigma_t, sigma_prev, style="base", norm_eps=False, refine_prior_sigma=False,
                           chunk_size=2):
        """(eps, eps_logvar, sigma_t, sigma_prev) as in src/experiments.py:400-460.  sigma_t / sigma_prev come back
        as [B,1,1,1] device tensors whenever the step made them per-sample (refine or 'pred*' styles)."""
        B = xt.shape[0]
        w = self._w(B)
        sch = self.scheduler
        sig_in = _as_f32(sigma_t, self.device)
        sp_in = _as_f32(sigma_prev, self.device)
        per_sample = refine_prior_sigma or "pred" in style or sig_in.numel() == B
        if refine_prior_sigma:
            ops.row_norm(xt, w.norms)
            ops.refine_sigma(w.norms, B, self.dim, sig_in, self.norm_min, self.norm_max, True, 0.0, sch.sigma_table,
                             self.time_shift, w.sigma, w.t, w.scale, slopes=sch.slopes_table)
        else:
            t_vec = torch.is_tensor(t) and t.numel() == B and B > 1
            ops.refine_sigma(None, B, self.dim, sig_in, 0.0, 0.0, False, 0.0 if t_vec else float(t), None, 0, w.sigma,
                             w.t, w.scale)
            if t_vec:
                w.t.copy_(t.reshape(-1).to(torch.float32).clamp_(0.0, 1000.0))
        sigma_cur, t_cur, scale_cur = w.sigma, w.t, w.scale
        sp_cur = sp_in
        forced = self.time_source(self._step) if self.time_source is not None else (None, None)
        if forced[0] is not None and refine_prior_sigma:
            w.t.copy_(forced[0].reshape(-1))
        if "pred" in style:
            feat = self.model.encode_scaled(xt, t_cur, scale_cur)
            r = self.sigma_model.forward_nhwc(feat)
            ops.sigma_correct(r, sigma_cur, sp_in, style == "pred", sch.sigma_table, w.sigma_hat, w.sigma_prev_hat,
                              w.t_hat, w.scale_hat, slopes=sch.slopes_table)
            if forced[1] is not None:
                w.t_hat.copy_(forced[1].reshape(-1))
            sigma_cur, t_cur, scale_cur = w.sigma_hat, w.t_hat, w.scale_hat
            sp_cur = w.sigma_prev_hat
        elif refine_prior_sigma and sp_in.numel() == 1:
            w.sigma_prev.copy_(sp_in.expand(B))
            sp_cur = w.sigma_prev
        out = self.model.forward_scaled(xt, t_cur, scale_cur)
        if self.learn_epsvar:
            C = out.shape[1] // 2
            w.eps.copy_(out[:, :C])
            learned = out[:, C:].contiguous()
        else:
            w.eps.copy_(out)
            learned = None
        if norm_eps:
            ops.normalize_rows_(w.eps)
        logvar = sch.get_eps_logvar(sigma_t=sigma_cur, sigma_prev=sp_cur, learned_logvar=learned)
        if per_sample:
            sp_ret = sp_cur.view(B, 1, 1, 1) if sp_cur.numel() == B else sp_cur
            return w.eps, logvar, sigma_cur.view(B, 1, 1, 1), sp_ret
        return w.eps, logvar, sigma_t, sigma_prev

    def _pred_xstart_clipped(self, xt, eps, sigma_t, out):
        """pred_xstart + clip_denoise_fn (src/schedulers.py:407-409, src/experiments.py:186-207)."""
        if self.clip_mode == CLIP_DYNAMIC:
            x0 = self.scheduler.pred_xstart(xt, eps, sigma_t, clip=CLIP_NONE, out=out)
            ops.dynamic_threshold_(x0, self.dynamic_ratio, self.dynamic_max)
            return x0
        return self.scheduler.pred_xstart(xt, eps, sigma_t, clip=self.clip_mode, out=out)

    # ---------------------------------------------------------------- L2: sigma feed-forward loop
    @torch.no_grad()
    def projection_loop(self, shape, gen=None, norm_init_noise=False, style="base", constrain_fn=None, norm_eps=False,
                        refine_prior_sigma=False, xT=None, return_log=False, chunk_size=2,
                        sigma_estimate_rate=(1, 0, 0, 0), constrain_loss=None, stop_condition=0.0, max_T=None,
                        sigma_pred_threshold=1000, new_eta=None, recal_sigma_prev=False, noise_fn=None, step_hook=None,
                        to_cpu=True, stop_check_every=1):
        """The module-level `projection_loop` of image_sample.py:431-519 (`--sampling project`): like denoise_loop but
        the next step's sigma is an estimate fed forward from this step,
            sigma <- r0*sigma_prev_orig + r1*sigma_prev + r2*sigma_t*||x_{t-1}||/||x_t|| + r3*dist(||x_{t-1}||),
        and its time is looked up from it (per sample).  The batch-global decisions of the reference (`t.max() >
        sigma_pred_threshold`, the mean constraint loss against `stop_condition`, :471,:514) read one scalar from the
        device per step, as the reference does; `stop_check_every` > 1 batches the stop test."""
        sch = self.scheduler
        sch.reset_state()
        sig = sch.sampling_sigmas
        T = sig.numel()
        if max_T is None:
            max_T = len(sch.timesteps_host) - 1
        if xT is None:
            xt, zt = self.get_noise_xt(shape=shape, gen=gen, norm_noise=norm_init_noise, sigma=sig[0])
        else:
            xt = xT
            zt = self.convert_coordinate(xt, sigma=sig[0]) if return_log else None
        B = xt.shape[0]
        w = self._w(B)
        w.nan_flag.zero_()
        w.xa.copy_(xt)
        xt, nxt = w.xa, w.xb
        f = lambda: torch.empty(B, device=self.device, dtype=torch.float32)
        last_norm, cur_norms, sig_est, t_est = f(), f(), f(), f()
        ops.row_norm(xt, last_norm)
        last_norm.div_(math.sqrt(self.dim))
        sigma_t = sig[0:1]
        t = sch.timesteps_host[0]
        t_max = float(t)
        sig_host = sig.cpu()
        rates = [float(r) for r in sigma_estimate_rate]
        z_list, eps_list, x0_prec_list, x0_postc_list, sigma_list, const_loss_list = [], [], [], [], [], []
        if return_log:
            z_list, sigma_list = [zt.cpu()], [sigma_t.cpu()]
        best_val, best_x0, x0, const_val = 10000, xt, xt, None
        steps = len(sch.timesteps_host)
        for ind in range(max_T):
            if ind == steps - 1 and new_eta is not None:
                sch.eta = new_eta
            sp_orig = sig[T - 1:T] if ind >= T - 1 else sig[ind + 1:ind + 2]
            if recal_sigma_prev:
                sigma_prev = _as_f32(sigma_t, self.device) * (sig[ind + 1] / sig[ind])
            else:
                sigma_prev = sp_orig
            cur_style, cur_refine = style, refine_prior_sigma
            if t_max > sigma_pred_threshold:
                cur_style, cur_refine = "base", False
            eps, eps_logvar, sigma_t, sigma_prev = self.get_denoise_vector(
                xt, t, sigma_t, sigma_prev, cur_style, norm_eps, refine_prior_sigma=cur_refine, chunk_size=chunk_size)
            x0_hat = self._pred_xstart_clipped(xt, eps, sigma_t, w.x0)
            x0 = constrain_fn(x0_hat) if constrain_fn is not None else x0_hat
            noise = noise_fn(ind, x0) if noise_fn is not None else None
            sch.pred_xprev(x0=x0, eps=eps, sigma_t=sigma_t, sigma_prev=sigma_prev, xt=xt, log_variance=eps_logvar,
                           noise=noise, out=nxt, nan_flag=w.nan_flag)
            ops.row_norm(nxt, cur_norms)
            ops.sigma_estimate(cur_norms, last_norm, self.dim, self.norm_max, float(sig_host[min(ind + 1, T - 1)]),
                               _as_f32(sigma_prev, self.device), _as_f32(sigma_t, self.device), rates, sch.sigma_table,
                               sch.slopes_table, sig_est, t_est)
            if step_hook is not None:
                step_hook(ind, dict(xt=xt, eps=eps, x0_hat=x0_hat, x0=x0, x_prev=nxt, sigma_t=sigma_t,
                                    sigma_prev=sigma_prev, sigma_next=sig_est, t_next=t_est))
            sigma_used, sigma_prev_used = sigma_t, sigma_prev
            sigma_t, t = sig_est.clone(), t_est.clone()
            t_max = float(t.max())  # one scalar D2H per step, as `t.max() > sigma_pred_threshold` in the reference
            if constrain_loss is not None:
                const, _ = constrain_loss(x0.clamp(-1, 1))
                const_val = torch.mean(const)
                if const_val < best_val:
                    best_x0, best_val = x0.clone(), const_val
                if return_log:
                    const_loss_list.append(const.cpu())
            else:
                best_x0 = x0
            if return_log:
                z_list.append(self.convert_coordinate(nxt, sigma=sigma_prev_used).cpu())
                eps_list.append(eps.cpu())
                x0_prec_list.append(x0_hat.cpu())
                x0_postc_list.append(x0.cpu())
                sigma_list.append(sigma_t.cpu())
            xt, nxt = nxt, xt
            if (ind + 1) % stop_check_every == 0:
                if int(w.nan_flag.item()) != 0 or (const_val is not None and float(const_val) <= stop_condition):
                    break
        result = best_x0.cpu() if to_cpu else best_x0
        return result, [z_list, eps_list, x0_prec_list, x0_postc_list, sigma_list, const_loss_list]

    # ---------------------------------------------------------------- L1: the DDIM-family loop
    @staticmethod
    def _device_loss(constrain_loss):
        """The device-resident twin of a `partial(Constraint_Function.loss, y=y)` (its `loss_device`), or None."""
        fn, kw = constrain_loss, {}
        if hasattr(fn, "func") and hasattr(fn, "keywords"):
            fn, kw = constrain_loss.func, dict(constrain_loss.keywords)
            if constrain_loss.args:
                return None
        owner = getattr(fn, "__self__", None)
        if owner is not None and getattr(fn, "__name__", "") == "loss" and hasattr(owner, "loss_device"):
            from functools import partial
            return partial(owner.loss_device, **kw)
        return None

    @staticmethod
    def _callable_signature(fn):
        """Identity of a `partial(Constraint_Function.method, y=y, ...)` for the graph cache: the bound object, the method and
        every keyword - tensors by (address, shape, dtype): a captured step reads them through their fixed addresses."""
        if fn is None:
            return None
        func, kw = (fn.func, fn.keywords) if hasattr(fn, "func") and hasattr(fn, "keywords") else (fn, {})
        if getattr(fn, "args", ()):
            return ("opaque", id(fn))
        items = tuple((k, (v.data_ptr(), tuple(v.shape), str(v.dtype)) if torch.is_tensor(v) else v) for k, v in sorted(kw.items()))
        return (id(getattr(func, "__self__", None)), getattr(func, "__name__", repr(func)), items)

    def _graph_signature(self, B, key, norm_eps, chunk_size, constrain_fn, loss_dev, world):
        """Everything a captured timestep bakes in besides device buffers with fixed addresses (the workspace of batch size B,
        the networks' plan buffers, the scheduler tables): the host scalars that become kernel arguments and the identities of
        the objects whose buffers it reads.  A cached graph is replayed only under an identical signature."""
        sch = self.scheduler
        return (B, key, bool(norm_eps), id(self.model), id(self.sigma_model), id(sch), sch.kind, float(sch.eta),
                getattr(sch, "sampler_var", None), self.clip_mode, float(self.norm_min), float(self.norm_max),
                bool(self.learn_epsvar), self.time_shift, self._callable_signature(constrain_fn),
                self._callable_signature(loss_dev), world, torch.cuda.current_stream().cuda_stream)

    def clear_graphs(self):
        """Drop the captured timesteps (they pin their private memory pools)."""
        self._graph_cache.clear()

    @torch.no_grad()
    def denoise_loop(self, shape, gen=None, norm_init_noise=False, style="base", constrain_fn=None, norm_eps=False,
                     refine_prior_sigma=False, xT=None, return_log=True, chunk_size=2, sigma_pred_threshold=1000,
                     new_eta=None, constrain_loss=None, return_best=True, free_const_steps=-1, noise_fn=None,
                     step_hook=None, to_cpu=True, graph=None, graph_skip=None, constrain_loss_device=None,
                     exact_global=False):
        """src/experiments.py:329-397.  `noise_fn(ind, like)` (optional) supplies the per-step noise instead of
        torch.randn_like (used by parity tests and by sharded runs that must reproduce the un-sharded stream);
        `step_hook(ind, dict)` (optional) observes per-step device tensors without copying them; `to_cpu=False`
        leaves the result on the device (the reference always returns a CPU tensor).

        `graph=True` (default: the experiment's `cuda_graph` attribute) replays the timestep from a CUDA graph: the
        ~300 kernel launches of one NLC step (refine -> encode -> sigma-model -> forward -> update [-> projection ->
        loss -> best-x0]) are captured once per loop call and style, and every later step costs three host calls (the
        step's two noise levels copied into a fixed buffer, the noise draw, the replay).  The arithmetic and the order of
        the random draws are those of the eager loop; the best-x0 decision moves to the device (nlc_best_update), which
        needs the device-resident twin of `constrain_loss` (`constrain_loss_device`, found automatically for
        Constraint_Function.loss).  Steps whose host scalars change (the 'base' steps above sigma_pred_threshold, a
        `new_eta` last step, steps without `refine_prior_sigma`, steps `graph_skip(ind)` excludes) run eagerly;
        `step_hook` sees the eager steps only.
        `exact_global=True` (sharded runs): the mean constraint loss behind the best-x0 choice and the NaN flag are
        all-reduced over the ranks, so every shard decides as the un-sharded reference does."""
        sch = self.scheduler
        sch.reset_state()
        sig = sch.sampling_sigmas
        if xT is None:
            xt, zt = self.get_noise_xt(shape=shape, gen=gen, norm_noise=norm_init_noise, sigma=sig[0])
        else:
            xt = xT
            zt = self.convert_coordinate(xt, sigma=sig[0]) if return_log else None
        B = xt.shape[0]
        w = self._w(B)
        w.nan_flag.zero_()
        w.xa.copy_(xt)
        xt, nxt = w.xa, w.xb
        eps_list, z_list, x0_prec_list, x0_postc_list, const_loss_list = [], [], [], [], []
        if return_log:
            z_list = [zt.cpu()]
        steps = sch.num_inference_steps
        ts_host = sch.timesteps_host.tolist()
        best_val, best_x0 = 10000, xt
        x0 = xt
        world = 1
        if exact_global:
            import torch.distributed as dist
            world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        use_graph = getattr(self, "cuda_graph", False) if graph is None else bool(graph)
        if self.time_source is not None or (step_hook is not None and graph_skip is None):
            use_graph = False  # (a forced time changes from step to step; a hook without `graph_skip` wants every step)
        loss_dev = constrain_loss_device
        if constrain_loss is not None and loss_dev is None and (use_graph or world > 1):
            loss_dev = self._device_loss(constrain_loss)
        # the device-side loop variant: fixed step buffers + device best-x0 tracking (graph replays and, for exactness
        # under sharding, the all-reduced loss); needs no per-step host read of the loss
        dev_mode = (use_graph or world > 1) and not return_log and (constrain_loss is None or loss_dev is not None)
        use_graph = use_graph and dev_mode and xt.is_cuda
        if dev_mode:
            gb = w.graph_bufs()
            gb.best_val.fill_(10000.0)
            gb.best_x0.copy_(xt)
            needs_noise = sch.kind in ("ddpm", "ddpm_orig") or float(sch.eta) > 0 or (new_eta is not None and new_eta > 0)
            # Captured timesteps live on the experiment across calls: re-capturing every pass costs a graph instantiation and,
            # when the old graph is destroyed, the cudaFree / cudaMalloc round trip of its private memory pool - 0.2 s per
            # pass, 10-20 % of a c2 pass (profiles/r02p_graph_cache.md)
            graphs, warmed = self._graph_cache, set()
            while len(graphs) > 8:
                graphs.pop(next(iter(graphs)))
            out_ref = {}

            def body(ind, t, sigma_t, sigma_prev, cur_style, cur_refine, apply_con, noise):
                eps, eps_logvar, s_t, s_p = self.get_denoise_vector(
                    xt, t, sigma_t, sigma_prev, cur_style, norm_eps, refine_prior_sigma=cur_refine, chunk_size=chunk_size)
                x0_hat = self._pred_xstart_clipped(xt, eps, s_t, w.x0)
                x0_ = constrain_fn(x0_hat) if apply_con else x0_hat
                sch.pred_xprev(x0=x0_, eps=eps, sigma_t=s_t, sigma_prev=s_p, xt=xt, log_variance=eps_logvar,
                               noise=noise, out=nxt, nan_flag=w.nan_flag)
                if loss_dev is not None:
                    const, _ = loss_dev(x0_.clamp(-1, 1))
                    torch.sum(const, dim=0, keepdim=True, out=gb.loss_sum)
                    if world > 1:
                        import torch.distributed as dist
                        dist.all_reduce(gb.loss_sum)
                    ops.best_update(gb.loss_sum, B * world, gb.best_val, gb.flag, x0_, gb.best_x0)
                xt.copy_(nxt)
                out_ref["x0"] = x0_
                return dict(xt=xt, eps=eps, x0_hat=x0_hat, x0=x0_, x_prev=nxt, sigma_t=s_t, sigma_prev=s_p)

            for ind in range(len(ts_host) - 1):
                t = ts_host[ind]
                self._step = ind
                last_eta = ind == steps - 1 and new_eta is not None
                if last_eta:
                    sch.eta = new_eta
                cur_style, cur_refine = style, refine_prior_sigma
                if t > sigma_pred_threshold:
                    cur_style, cur_refine = "base", False
                apply_con = constrain_fn is not None and (free_const_steps <= 0 or ind <= free_const_steps)
                gb.step_sig.copy_(sig[ind:ind + 2], non_blocking=True)
                noise = None
                if needs_noise and (sch.kind in ("ddpm", "ddpm_orig") or float(sch.eta) > 0):
                    if noise_fn is not None:
                        gb.noise.copy_(noise_fn(ind, xt))
                    else:
                        gb.noise.normal_()  # the draw torch.randn_like(x0) makes, on the same generator
                    noise = gb.noise
                key = (cur_style, apply_con, noise is not None)
                capturable = (use_graph and cur_refine and not last_eta
                              and (graph_skip is None or not graph_skip(ind)))
                gkey = self._graph_signature(B, key, norm_eps, chunk_size, constrain_fn if apply_con else None, loss_dev,
                                             world) if capturable else None
                if capturable and (key in warmed or gkey in graphs):
                    g = graphs.get(gkey)
                    if g is None:
                        n0 = ops.STATS.launches
                        g = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(g):
                            body(ind, t, gb.step_sig[0:1], gb.step_sig[1:2], cur_style, cur_refine, apply_con, noise)
                        g.n_launches = ops.STATS.launches - n0
                        g.x0 = out_ref["x0"]  # (lives in the graph's memory pool: every replay refreshes it)
                        ops.STATS.launches = n0
                        graphs[gkey] = g
                    g.replay()
                    out_ref["x0"] = g.x0
                    ops.STATS.launches += g.n_launches
                    ops.STATS.graph_replays += 1
                    sch.i += 1
                else:
                    rec = body(ind, t, gb.step_sig[0:1], gb.step_sig[1:2], cur_style, cur_refine, apply_con, noise)
                    if cur_refine:
                        warmed.add(key)
                    if step_hook is not None:
                        step_hook(ind, rec)
                if (ind + 1) % self.nan_check_every == 0:
                    if world > 1:
                        import torch.distributed as dist
                        flag = w.nan_flag.clone()
                        dist.all_reduce(flag, op=dist.ReduceOp.MAX)
                    else:
                        flag = w.nan_flag
                    if int(flag.item()) != 0:
                        break
            x0 = out_ref.get("x0", xt)
            best_x0 = gb.best_x0 if loss_dev is not None else x0
            result = (best_x0 if return_best else x0).clone()
            result = result.cpu() if to_cpu else result
            return result, [z_list, eps_list, x0_prec_list, x0_postc_list, const_loss_list]

        for ind in range(len(ts_host) - 1):
            t = ts_host[ind]
            self._step = ind
            if ind == steps - 1 and new_eta is not None:
                sch.eta = new_eta
            sigma_t, sigma_prev = sig[ind:ind + 1], sig[ind + 1:ind + 2]
            cur_style, cur_refine = style, refine_prior_sigma
            if t > sigma_pred_threshold:
                cur_style, cur_refine = "base", False
            eps, eps_logvar, sigma_t, sigma_prev = self.get_denoise_vector(
                xt, t, sigma_t, sigma_prev, cur_style, norm_eps, refine_prior_sigma=cur_refine, chunk_size=chunk_size)
            x0_hat = self._pred_xstart_clipped(xt, eps, sigma_t, w.x0)
            if constrain_fn is not None and (free_const_steps <= 0 or ind <= free_const_steps):
                x0 = constrain_fn(x0_hat)
            else:
                x0 = x0_hat
            noise = noise_fn(ind, x0) if noise_fn is not None else None
            sch.pred_xprev(x0=x0, eps=eps, sigma_t=sigma_t, sigma_prev=sigma_prev, xt=xt, log_variance=eps_logvar,
                           noise=noise, out=nxt, nan_flag=w.nan_flag)
            if step_hook is not None:
                step_hook(ind, dict(xt=xt, eps=eps, x0_hat=x0_hat, x0=x0, x_prev=nxt, sigma_t=sigma_t,
                                    sigma_prev=sigma_prev))
            if constrain_loss is not None:
                const, _ = constrain_loss(x0.clamp(-1, 1))
                const_val = torch.mean(const)
                if const_val < best_val:
                    best_x0 = x0.clone()
                    best_val = const_val
                if return_log:
                    const_loss_list.append(const.cpu())
            else:
                best_x0 = x0
            if return_log:
                z_list.append(self.convert_coordinate(nxt, sigma=sigma_prev).cpu())
                eps_list.append(eps.cpu())
                x0_prec_list.append(x0_hat.cpu())
                x0_postc_list.append(x0.cpu())
            xt, nxt = nxt, xt
            if (ind + 1) % self.nan_check_every == 0 and int(w.nan_flag.item()) != 0:
                break
        result = best_x0 if return_best else x0
        result = result.cpu() if to_cpu else result
        return result, [z_list, eps_list, x0_prec_list, x0_postc_list, const_loss_list]


class ImageExperiment(ExperimentDiffusion):
    """src/experiments.py:553-559."""

    def __init__(self, model, scheduler, batch_size=64, data_shape=(3, 32, 32), seed=0, device="cuda:0",
                 save_folder="./", dist_train=False, time_shift=0):
        super().__init__(model=model, scheduler=scheduler, batch_size=batch_size, data_shape=data_shape,
                         save_folder=save_folder, seed=seed, device=device, dist_train=dist_train,
                         time_shift=time_shift)


class StackedRandomGenerator:
    """src/experiments.py:71-85: one torch.Generator per sample, seeded by the sample's global index, so a batch
    (or a shard of it) draws the same latents wherever it runs."""

    def __init__(self, device, seeds):
        self.generators = [torch.Generator(device).manual_seed(int(seed) % (1 << 32)) for seed in seeds]

    def randn(self, size, **kwargs):
        assert size[0] == len(self.generators)
        return torch.stack([torch.randn(size[1:], generator=gen, **kwargs) for gen in self.generators])

    def randn_like(self, input):
        return self.randn(input.shape, dtype=input.dtype, layout=input.layout, device=input.device)


class _EdmWork:
    def __init__(self, B, shape, device):
        f32 = lambda *s: torch.empty(*s, device=device, dtype=torch.float32)
        f64 = lambda *s: torch.empty(*s, device=device, dtype=torch.float64)
        from ._lib import EDM_PARTS
        self.x32 = f32(B, *shape)
        self.eps_raw, self.eps, self.eps_next, self.denoised = (f64(B, *shape) for _ in range(4))
        self.e1, self.ne = f64(B, *shape), f64(B, *shape)
        self.xa, self.xb, self.xh = f64(B, *shape), f64(B, *shape), f64(B, *shape)
        self.parts = f64(B, EDM_PARTS)
        self.parts3 = f64(B, EDM_PARTS, 3)


def _is_single(t):
    """`len(t.unsqueeze(-1)) == 1` of the reference (src/experiments.py:812-832)."""
    return t.dim() == 0 or t.shape[0] == 1


class EDMImageExperiment(ImageExperiment):
    """src/experiments.py:756-961 (sampling methods): EDM preconditioning, NLC with the EDM sigma-model, Heun sampler.

    The sample stays float64 on the device; every O(B*d) operation is a libnlc_b200 kernel (nlc_edm_*).  The
    per-sample noise levels ([B] or scalar tensors) follow the reference's expressions literally, which also
    reproduces its float32 / float64 type promotion (a 0-d float64 sigma times a float32 [B,1,1,1] correction is
    float32, src/experiments.py:824)."""

    def __init__(self, model, scheduler, batch_size=64, data_shape=(3, 32, 32), seed=0, device="cuda:0",
                 save_folder="./", dist_train=False, time_shift=0, sigma_min=0.002, sigma_max=80, rho=7, S_churn=0,
                 S_min=0, S_max=float("inf"), S_noise=1, sigma_data=0.5, P_mean=-1.2, P_std=1.2, num_timesteps=18):
        super().__init__(model=model, scheduler=scheduler, batch_size=batch_size, data_shape=data_shape, seed=seed,
                         device=device, save_folder=save_folder, dist_train=dist_train, time_shift=time_shift)
        self.sigma_min, self.sigma_max, self.rho = sigma_min, sigma_max, rho
        self.S_churn, self.S_min, self.S_max, self.S_noise = S_churn, S_min, S_max, S_noise
        self.sigma_data, self.P_mean, self.P_std, self.num_timesteps = sigma_data, P_mean, P_std, num_timesteps
        self._ework = {}

    def _ew(self, B):
        w = self._ework.get(B)
        if w is None:
            w = _EdmWork(B, self.data_shape, self.device)
            self._ework[B] = w
        return w

    def _precond(self, sigma, B):
        """c_skip, c_out, c_in, c_noise as [B] float32 (src/experiments.py:794-797)."""
        sigma = sigma.to(self.device, torch.float32).reshape(-1)
        sd = self.sigma_data
        c_skip = sd ** 2 / (sigma ** 2 + sd ** 2)
        c_out = sigma * sd / (sigma ** 2 + sd ** 2).sqrt()
        c_in = 1 / (sd ** 2 + sigma ** 2).sqrt()
        c_noise = sigma.log() / 4
        return [c.expand(B).contiguous() for c in (c_skip, c_out, c_in, c_noise)]

    @torch.no_grad()
    def get_denoise_vector(self, xt, sigma_t, sigma_prev, style="base", norm_eps=False, refine_prior_sigma=False,
                           _eps_out=None):
        """(eps, denoised, sigma_t, sigma_prev) as in src/experiments.py:805-843; xt float64 [B,C,H,W]."""
        B = xt.shape[0]
        w = self._ew(B)
        dev = self.device
        sigma_t, sigma_prev = torch.as_tensor(sigma_t), torch.as_tensor(sigma_prev)
        sigma_t_orig = sigma_t
        xt = xt.contiguous()
        ops.edm_prepare(xt, w.x32, w.parts if refine_prior_sigma else None)
        if refine_prior_sigma:
            norm_x = w.parts.sum(1).sqrt().view(B, 1, 1, 1) / math.sqrt(self.dim)
            min_dist = torch.clamp(norm_x - self.norm_max, min=0)
            max_dist = norm_x + self.norm_min
            raw_sigma = torch.ones_like(norm_x) * sigma_t if _is_single(sigma_t) else sigma_t
            sigma_t = torch.clamp(raw_sigma, min=min_dist, max=max_dist)
            if _is_single(sigma_prev):
                sigma_prev = torch.ones_like(norm_x) * sigma_prev
        if "pred" in style:
            _, _, c_in, c_noise = self._precond(sigma_t, B)
            feat = self.model.encode_scaled(w.x32, c_noise, c_in)
            r = self.sigma_model.forward_nhwc(feat).view(B, 1, 1, 1)
            dist_hat = sigma_t * (1 + r)
            dist_prev_hat = dist_hat * (sigma_prev / sigma_t)
            sigma_t = dist_hat
            if style == "pred":
                sigma_prev = dist_prev_hat
        # (dimensioned host tensors cannot meet device tensors; moving them does not change the dtype rules)
        if _is_single(sigma_t_orig):
            sigma_t_orig = sigma_t_orig.reshape(-1, 1, 1, 1).to(dev)
        if _is_single(sigma_t):
            sigma_t = sigma_t.reshape(-1, 1, 1, 1).to(dev)
        if _is_single(sigma_prev):
            sigma_prev = sigma_prev.reshape(-1, 1, 1, 1).to(dev)
        used = sigma_t_orig if style == "pred_sigma" else sigma_t
        c_skip, c_out, c_in, c_noise = self._precond(used, B)
        F_x = self.model.forward_scaled(w.x32, c_noise, c_in)
        div = used.to(dev, torch.float64).reshape(-1).expand(B).contiguous()
        eps = _eps_out if _eps_out is not None else w.eps
        if norm_eps:
            ops.edm_eps(xt, w.x32, F_x, c_skip, c_out, div, w.eps_raw, w.denoised, w.parts)
            den = torch.clamp(w.parts.sum(1).sqrt(), min=1e-12)
            ops.edm_mix(w.eps_raw, den, None, None, None, None, 1.0, 0.0, eps)
        else:
            ops.edm_eps(xt, w.x32, F_x, c_skip, c_out, div, eps, w.denoised, None)
        return eps, w.denoised, sigma_t, sigma_prev

    def _vec64(self, v, B):
        return torch.as_tensor(v).to(self.device, torch.float64).reshape(-1).expand(B).contiguous()

    @torch.no_grad()
    def edm_sampler(self, shape, gen=None, style="base,base", norm_eps="000", refine_prior_sigma=False, num_steps=None,
                    sigma_scheduler="EDM", eps_ratio=0.5, eps_scale=1.0, use_second_order=True, latents=None,
                    step_hook=None):
        """src/experiments.py:847-918.  `latents` (optional) replaces gen.randn (parity tests, sharded runs);
        `step_hook(i, dict)` observes per-step device tensors."""
        norm_eps, norm_eps_combine = bool(int(norm_eps[0])), bool(int(norm_eps[1]))
        style_t, style_next = style.split(",")
        num_steps = self.num_timesteps if num_steps is None else num_steps
        if latents is None:
            latents = gen.randn(shape, device=self.device)
        B = latents.shape[0]
        w = self._ew(B)
        # noise levels stay on the host (0-d float64 tensors): no device->host read in the loop
        step_indices = torch.arange(num_steps, dtype=torch.float64)
        if sigma_scheduler == "EDM":
            sigma_steps = (self.sigma_max ** (1 / self.rho) + step_indices / (num_steps - 1) * (
                self.sigma_min ** (1 / self.rho) - self.sigma_max ** (1 / self.rho))) ** self.rho
        elif sigma_scheduler == "Linear":
            sigma_steps = torch.tensor(np.exp(np.linspace(np.log(self.sigma_max), np.log(self.sigma_min), num_steps)))
        else:
            raise NotImplementedError
        sigma_steps = torch.cat([torch.as_tensor(sigma_steps), torch.zeros_like(sigma_steps[:1])])
        w.xa.copy_(latents.to(torch.float64) * sigma_steps[0])
        x_next, x_spare = w.xa, w.xb
        for i, (sigma_cur, sigma_next) in enumerate(zip(sigma_steps[:-1], sigma_steps[1:])):
            x_cur = x_next
            sigma_next0 = sigma_next
            gamma = min(self.S_churn / num_steps, np.sqrt(2) - 1) if self.S_min <= sigma_cur <= self.S_max else 0
            sigma_hat = torch.as_tensor(sigma_cur + gamma * sigma_cur)
            sigma_hat0 = sigma_hat
            if gamma > 0:
                z = torch.randn_like(x_cur)
                ops.edm_axpy(x_cur, z, None, 0.0, None,
                             self._vec64((sigma_hat ** 2 - sigma_cur ** 2).sqrt() * self.S_noise, B), w.xh)
                x_hat = w.xh
            else:
                x_hat = x_cur
            eps, denoised, sigma_hat, sigma_next = self.get_denoise_vector(
                x_hat, sigma_hat, sigma_next, style=style_t, norm_eps=norm_eps, refine_prior_sigma=refine_prior_sigma)
            # eps <- eps * (sigma_hat / sigma_hat0)
            ops.edm_mix(eps, None, self._vec64(sigma_hat / sigma_hat0, B), None, None, None, 1.0, 0.0, w.e1)
            if "pred_partial" in style_t:
                sigma_next = sigma_next0
            if style_t == "pred_partial":
                coef = sigma_next - sigma_hat0
            else:
                coef = sigma_next - sigma_hat
            ops.edm_axpy(x_hat, w.e1, None, 0.0, None, self._vec64(coef, B), x_spare)
            if style_t == "pred_partial3":
                sigma_hat = sigma_hat0
            if i < num_steps - 1 and use_second_order:
                eps_next, denoised, sigma_next, _ = self.get_denoise_vector(
                    x_spare, sigma_next, sigma_next * 0, style=style_next, norm_eps=norm_eps,
                    refine_prior_sigma=refine_prior_sigma, _eps_out=w.eps_next)
                s2 = self._vec64(sigma_next / sigma_next0, B)
                if "pred_partial" in style_next:
                    sigma_next = sigma_next0
                # new_eps = eps_ratio * eps + (1 - eps_ratio) * eps_next * (sigma_next / sigma_next0)
                ops.edm_mix(w.e1, None, None, eps_next, None, s2, eps_ratio, 1 - eps_ratio, w.ne, w.parts3)
                den = mul = None
                if norm_eps_combine:
                    den = torch.clamp(w.parts3[:, :, 0].sum(1).sqrt(), min=1e-12)
                if eps_scale is None:
                    # CosineSimilarity(dim=1, eps=1e-6) between new_eps and eps (scale invariant, so the optional
                    # normalisation of new_eps does not change it)
                    s = w.parts3.sum(1)
                    mul = s[:, 2] / (s[:, 0].sqrt().clamp_min(1e-6) * s[:, 1].sqrt().clamp_min(1e-6))
                    mul = mul.contiguous()
                ops.edm_axpy(x_hat, w.ne, den, eps_scale, mul, self._vec64(sigma_next - sigma_hat, B), x_spare)
            if step_hook is not None:
                step_hook(i, dict(x_hat=x_hat, x_next=x_spare, sigma_hat=sigma_hat, sigma_next=sigma_next, eps=w.e1))
            x_next, x_spare = x_spare, x_next
        return x_next

    @torch.no_grad()
    def evaluate_edm(self, n_samples, images_dir=None, gen=None, style="base,base", norm_eps="000",
                     refine_prior_sigma=False, microbatch=-1, sigma_scheduler="EDM", eps_ratio=0.5, eps_scale=1.0,
                     use_second_order=True, rank=0, world=1):
        """Sampling part of src/experiments.py:923-961: per-sample seeds arange(n_samples) split into batches
        (:929-933); batches are dealt round-robin to `world` ranks.  Returns the images in [0,1]
        (`sample.add(1).div(2).clamp(0,1)`, :946); PNG writing and FID are outside the hot path."""
        batch_size = microbatch if microbatch > 0 else self.batch_size
        seeds = np.arange(n_samples)
        num_batches = (len(seeds) - 1) // batch_size + 1
        all_batches = torch.as_tensor(seeds).tensor_split(num_batches)
        out = []
        for i in range(rank, num_batches, world):
            g = StackedRandomGenerator(self.device, all_batches[i])
            shape = (len(all_batches[i]),) + self.data_shape
            sample = self.edm_sampler(shape=shape, gen=g, style=style, norm_eps=norm_eps,
                                      refine_prior_sigma=refine_prior_sigma, sigma_scheduler=sigma_scheduler,
                                      eps_ratio=eps_ratio, eps_scale=eps_scale, use_second_order=use_second_order)
            out.append(sample.add(1).div(2).clamp(0, 1))
        return {"samples": torch.cat(out) if out else None}
