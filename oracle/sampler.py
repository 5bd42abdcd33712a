"""ORACLE (test infrastructure, not product code) — CPU fp32 restatement of the reference's sampler arithmetic:
sigma tables and timestep selection, the NLC step front-end and the DDIM-family update loop.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
Pinned against the reference itself: tests/test_oracle_vs_reference.py (live, when /root/reference exists) and
tests/golden/*.pt (per-step dumps of the reference's own ImageExperiment.denoise_loop).

Follows src/schedulers.py:95-164 (tables), :227-284 (timestep selection), :367-390 (get_eps_logvar), :407-409
(pred_xstart), :432-627 (pred_xprev variants); src/experiments.py:263-293 (noise / coordinates), :329-397
(denoise_loop), :400-460 (get_denoise_vector); src/utils.py:7-16 (vector_norm, normalize).
"""
import math

import numpy as np
import torch


# ------------------------------------------------------------------------------------------------ tables
class Tables:
    def __init__(self, n_train=1000, beta_start=0.0001, beta_end=0.02):
        self.betas = torch.linspace(beta_start, beta_end, n_train, dtype=torch.float32)
        self.alphas_cumprod = torch.cumprod(1.0 - self.betas, dim=0)
        self.sigmas = (1 / self.alphas_cumprod - 1).sqrt()  # src/schedulers.py:134
        prev = torch.cat([torch.ones(1), self.alphas_cumprod[:-1]])
        self.posterior_variance = self.betas * (1.0 - prev) / (1.0 - self.alphas_cumprod)

    continuous = False  # continuous_t (src/schedulers.py:232): Interp1d time lookup instead of searchsorted

    @staticmethod
    def _interp(x, y, xnew):
        """src/torchinterp1d.py:96-148 for 1-D x, y."""
        eps = torch.finfo(y.dtype).eps
        ind = torch.clamp(torch.searchsorted(x.contiguous(), xnew.contiguous()) - 1, 0, x.shape[0] - 2)
        slopes = (y[1:] - y[:-1]) / (eps + (x[1:] - x[:-1]))
        return y[ind] + slopes[ind] * (xnew - x[ind])

    def t_of_sigma(self, sigma):
        if self.continuous:  # sigma_to_t_interp (:210-220)
            shape = sigma.shape
            xnew = sigma.squeeze()
            xnew = xnew.unsqueeze(0) if xnew.dim() == 0 else xnew
            t = self._interp(self.sigmas, torch.arange(len(self.sigmas)).float(), xnew.reshape(-1)).float()
            return t.reshape(xnew.shape) if len(shape) else t
        return torch.searchsorted(self.sigmas, sigma)  # first index with table >= sigma (:185-190)

    def sigma_of_t(self, t):
        """t_to_sigma_interp (:192-203)."""
        y = self._interp(torch.arange(len(self.sigmas)).float(), self.alphas_cumprod, t)
        return torch.where(t >= 0, (1 / y - 1).sqrt(), torch.zeros(())).float()

    def ddim_schedule(self, start_sigma, end_sigma, n_steps):
        """Timesteps and sigmas of style 'DDIM' with set_alpha_to_one (src/schedulers.py:237-284)."""
        start = torch.as_tensor(min(float(start_sigma), float(self.sigmas[-1])), dtype=torch.float32) \
            if start_sigma is not None else self.sigmas[-1]
        end = self.sigmas[0] if end_sigma is None else torch.as_tensor(end_sigma, dtype=torch.float32)
        if self.continuous:
            t_hi, t_lo = self.t_of_sigma(start).item(), self.t_of_sigma(end).item()
            span = t_hi + 1 - t_lo
            if span != int(span):
                span = span // 1 + 1  # space_timesteps on a fractional count (:73-77)
        else:
            t_hi, t_lo = int(self.t_of_sigma(start)), int(self.t_of_sigma(end))
            span = t_hi + 1 - t_lo
        stride = 1 if n_steps <= 1 else (span - 1) / (n_steps - 1)
        picks, cur = set(), 0.0
        for _ in range(n_steps):
            picks.add(round(cur))
            cur += stride
        if self.continuous:
            ts = torch.tensor(t_lo + np.array(sorted(picks, reverse=True)), dtype=torch.float32)
            sig = self.sigma_of_t(ts)
            ts = torch.cat([ts, torch.tensor([-1])])
            sig = torch.cat([sig, torch.zeros(1)])
            s_t, s_p = sig[-3], sig[-2]
            beta_t = (s_t ** 2 - s_p ** 2) / (s_t ** 2 + 1)
            return ts, sig, beta_t * (1 - 1 / (s_p ** 2 + 1)) / (1 - 1 / (s_t ** 2 + 1))
        ts = [t_lo + v for v in sorted(picks, reverse=True)]
        # strictly decreasing repair (:15-31)
        n = len(ts)
        up = [0] * n
        up[-2:] = ts[-2:]
        for i in range(n - 1, 0, -1):
            up[i - 1] = ts[i - 1] if ts[i - 1] > up[i] else up[i] + 1
        fixed, cap = [0] * n, 999
        for i in range(n - 1):
            fixed[i] = min(up[i], cap)
            cap = fixed[i] - 1
        ts = torch.tensor(fixed, dtype=torch.long)
        sig = self.sigmas[ts]
        ts = torch.cat([ts, torch.tensor([-1])])
        sig = torch.cat([sig, torch.zeros(1)])
        s_t, s_p = sig[-3], sig[-2]
        beta_t = (s_t ** 2 - s_p ** 2) / (s_t ** 2 + 1)
        min_var_coef = beta_t * (1 - 1 / (s_p ** 2 + 1)) / (1 - 1 / (s_t ** 2 + 1))
        return ts, sig, min_var_coef


# ------------------------------------------------------------------------------------------------ step arithmetic
def vector_norm(x):
    return torch.linalg.vector_norm(x, dim=tuple(range(1, x.dim())), keepdim=True)


def normalize(x, dim):
    return math.sqrt(dim) * x / torch.clamp(vector_norm(x), min=1e-12)


def eps_logvar(sigma_t, sigma_prev, min_var_coef, mode, learned=None):
    """src/schedulers.py:367-390."""
    if mode == "none":
        return None
    beta_t = ((sigma_t ** 2 - sigma_prev ** 2) / (sigma_t ** 2 + 1)).abs().clamp(min=1e-20)
    a_t, a_p = 1 / (sigma_t ** 2 + 1), 1 / (sigma_prev ** 2 + 1)
    coef = ((1 - a_p) / (1 - a_t)).clamp(min=0, max=1)
    max_lv = beta_t.log()
    min_lv = (beta_t * coef).clamp(min=min_var_coef).log()
    if mode == "learned":
        frac = (learned + 1) / 2
        return frac * max_lv + (1 - frac) * min_lv
    return min_lv if mode == "fixedsmall" else max_lv


def pred_xprev(kind, eta, x0, eps, sigma_t, sigma_prev, xt, logvar, noise):
    """All pred_xprev variants, src/schedulers.py:432-627; `noise` stands for the torch.randn_like draw."""
    if kind in ("ddim_simple_orig", "ddim_simple_drag", "ddim_orig"):
        eps = (xt - x0) / sigma_t
    if kind in ("ddim", "ddim_orig"):
        if eta > 0:
            noise_sigma = eta * torch.exp(0.5 * logvar) / torch.sqrt(1 / (sigma_prev ** 2 + 1))
            nz = (sigma_prev > 0) * noise
        else:
            noise_sigma, nz = 0, 0
        signal = torch.sqrt((sigma_prev ** 2 - noise_sigma ** 2).clamp(min=0))
        if kind == "ddim":
            noise_sigma = torch.sqrt(sigma_prev ** 2 - signal ** 2)
        return x0 + signal * eps + noise_sigma * nz
    if kind in ("ddim_simple", "ddim_simple_orig", "ddim_simple_drag"):
        signal = sigma_prev if kind == "ddim_simple_drag" else math.sqrt(1 - eta ** 2) * sigma_prev
        out = x0 + signal * eps
        if eta > 0:
            out = out + (eta * sigma_prev) * noise
        return out
    if kind == "ddpm":
        noise_sigma = torch.exp(0.5 * logvar) / torch.sqrt(1 / (sigma_prev ** 2 + 1))
        signal = torch.sqrt((sigma_prev ** 2 - noise_sigma ** 2).clamp(min=0))
        return x0 + signal * eps + noise_sigma * ((sigma_prev > 0) * noise)
    if kind == "ddpm_orig":
        ab, abp = 1 / (sigma_t ** 2 + 1), 1 / (sigma_prev ** 2 + 1)
        a_t = ab / abp
        mean = (1 - a_t) * abp.sqrt() / (1.0 - ab) * x0 + (1.0 - abp) * a_t.sqrt() / (1.0 - ab) * (xt * ab.sqrt())
        zprev = mean + (sigma_prev > 0).float() * torch.exp(0.5 * logvar) * noise
        return zprev / abp.sqrt()
    raise ValueError(kind)


def denoise_vector(tab, model_fwd, model_enc, sigma_fn, xt, t, sigma_t, sigma_prev, style, norm_eps, refine,
                   norm_min, norm_max, time_shift=0, learn_epsvar=False):
    """src/experiments.py:400-460.  model_fwd/model_enc take (z, t) with z = x/sqrt(sigma^2+1)."""
    dim = xt[0].numel()
    B = xt.shape[0]
    if refine:
        nx = vector_norm(xt) / math.sqrt(dim)
        lo, hi = torch.clamp(nx - norm_max, min=0), nx + norm_min
        sigma_t = torch.clamp(torch.ones_like(nx) * sigma_t, min=lo, max=hi)
        t = tab.t_of_sigma(sigma_t)
        if t.min() > 0:
            t = t - time_shift
        sigma_prev = torch.ones_like(nx) * sigma_prev
    t = torch.clamp(torch.as_tensor(t), min=0.0, max=1000.0)

    def scaled(x, s):
        return x * (1 / (s ** 2 + 1)).sqrt()

    def batched(tt):
        tt = tt.reshape(-1).float()
        return tt if tt.numel() == B else torch.ones(B) * tt

    if "pred" in style:
        feat = model_enc(scaled(xt, sigma_t), batched(t))
        r = sigma_fn(feat)
        dist = sigma_t * (1 + r)
        dist_prev = dist * (sigma_prev / sigma_t)
        t = torch.clamp(tab.t_of_sigma(dist), min=0.0, max=1000.0)
        sigma_t = dist
        if style == "pred":
            sigma_prev = dist_prev
    out = model_fwd(scaled(xt, sigma_t), batched(t))
    learned = None
    if learn_epsvar:
        C = out.shape[1] // 2
        out, learned = out[:, :C], out[:, C:]
    if norm_eps:
        out = normalize(out, dim)
    return out, learned, sigma_t, sigma_prev


def clip_x0(x0_hat, clip):
    """clip_denoise_fn (src/experiments.py:186-207)."""
    if clip == "clamp":
        return x0_hat.clamp(-1, 1)
    if clip == "dynamic":
        b = x0_hat.shape[0]
        s = torch.quantile(x0_hat.reshape(b, -1).abs(), 0.99, dim=1).clamp(min=1, max=100).view(b, 1, 1, 1)
        return torch.clamp(x0_hat, -s, s) / s
    return x0_hat


def projection_loop(tab, ts, sig, min_var_coef, model_fwd, model_enc, sigma_fn, xT, kind="ddim", eta=0.0,
                    sampler_var="none", style="pred", norm_eps=True, refine=True, norm_min=0.0, norm_max=1.0,
                    clip="clamp", constrain_fn=None, noises=None, sigma_pred_threshold=1000, rates=(1, 0, 0, 0),
                    max_T=None, recal_sigma_prev=False, log=None):
    """The module-level projection_loop of image_sample.py:431-519 (no constraint loss / early stop)."""
    dim = xT[0].numel()
    xt = x0 = xT
    sigma_t, t = sig[0], ts[0]
    last_norm = vector_norm(xt) / math.sqrt(dim)
    T = len(sig)
    max_T = len(ts) - 1 if max_T is None else max_T
    for ind in range(max_T):
        sp_orig = sig[-1] if ind >= T - 1 else sig[ind + 1]
        sigma_prev = sigma_t * (sig[ind + 1] / sig[ind]) if recal_sigma_prev else sp_orig
        cur_style, cur_refine = (style, refine) if not (torch.as_tensor(t).max() > sigma_pred_threshold) else ("base", False)
        eps, learned, sigma_t, sigma_prev = denoise_vector(tab, model_fwd, model_enc, sigma_fn, xt, t, sigma_t,
                                                           sigma_prev, cur_style, norm_eps, cur_refine, norm_min,
                                                           norm_max)
        logvar = eps_logvar(sigma_t, sigma_prev, min_var_coef, sampler_var, learned)
        x0_hat = clip_x0(xt - sigma_t * eps, clip)
        x0 = constrain_fn(x0_hat) if constrain_fn is not None else x0_hat
        x_prev = pred_xprev(kind, eta, x0, eps, sigma_t, sigma_prev, xt, logvar,
                            noises[ind] if noises is not None else None)
        cur_norm = vector_norm(x_prev) / math.sqrt(dim)
        cur_dist = torch.sqrt(cur_norm ** 2 + norm_max ** 2 - 2 * cur_norm * norm_max * 0.99 + 1e-8)
        new_sigma = rates[0] * sp_orig + rates[1] * sigma_prev + rates[2] * (sigma_t * (cur_norm / last_norm)) \
            + rates[3] * cur_dist
        if log is not None:
            log.append(dict(xt=xt, eps=eps, x0=x0, x_prev=x_prev, sigma_t=torch.as_tensor(sigma_t).reshape(-1),
                            sigma_next=new_sigma.reshape(-1)))
        sigma_t, t, last_norm, xt = new_sigma, tab.t_of_sigma(new_sigma), cur_norm, x_prev
    return x0


def denoise_loop(tab, ts, sig, min_var_coef, model_fwd, model_enc, sigma_fn, xT, kind="ddim", eta=0.0,
                 sampler_var="none", style="pred", norm_eps=True, refine=True, norm_min=0.0, norm_max=1.0, clip="clamp",
                 constrain_fn=None, noises=None, sigma_pred_threshold=1000, learn_epsvar=False, log=None,
                 constrain_loss=None):
    """src/experiments.py:329-397 with the clip functions of :186-207.  `noises[ind]` replaces torch.randn_like.
    `log` (a list) receives one dict of tensors per step.  With `constrain_loss` the x0 of the step with the lowest
    batch-mean constraint loss is returned (best-x0 tracking, :371-381); otherwise the last x0."""
    xt = xT
    x0 = xt
    best_val, best_x0 = 10000, xt
    for ind in range(len(ts) - 1):
        t = ts[ind]
        sigma_t, sigma_prev = sig[ind], sig[ind + 1]
        cur_style, cur_refine = (style, refine) if t <= sigma_pred_threshold else ("base", False)
        eps, learned, sigma_t, sigma_prev = denoise_vector(tab, model_fwd, model_enc, sigma_fn, xt, t, sigma_t,
                                                           sigma_prev, cur_style, norm_eps, cur_refine, norm_min,
                                                           norm_max, learn_epsvar=learn_epsvar)
        logvar = eps_logvar(sigma_t, sigma_prev, min_var_coef, sampler_var, learned)
        x0_hat = clip_x0(xt - sigma_t * eps, clip)
        x0 = constrain_fn(x0_hat) if constrain_fn is not None else x0_hat
        noise = noises[ind] if noises is not None else None
        x_prev = pred_xprev(kind, eta, x0, eps, sigma_t, sigma_prev, xt, logvar, noise)
        const = None
        if constrain_loss is not None:
            const, _ = constrain_loss(x0.clamp(-1, 1))
            if torch.mean(const) < best_val:
                best_x0, best_val = x0.clone(), torch.mean(const)
        else:
            best_x0 = x0
        if log is not None:
            log.append(dict(xt=xt, eps=eps, x0_hat=x0_hat, x0=x0, x_prev=x_prev, const=const,
                            sigma_t=torch.as_tensor(sigma_t).reshape(-1), sigma_prev=torch.as_tensor(sigma_prev).reshape(-1)))
        xt = x_prev
    return best_x0
