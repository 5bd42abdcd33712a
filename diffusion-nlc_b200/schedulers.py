"""Host-side mirror of the reference's sigma-parameterised schedulers (src/schedulers.py).

Table construction (betas, cumulative alphas, sigma_t = sqrt(1/alpha_bar_t - 1), timestep selection) is one-time
host work done with the same torch CPU operations as the reference so the tables are bit-identical; every
per-step tensor operation (`pred_xstart`, `pred_xprev`, `get_eps_logvar`, `get_t_from_sigma`) is a call into
libnlc_b200.  `get_sampler` keeps the reference's signature (src/schedulers.py:676-726).
"""
import numpy as np
import torch

from . import ops
from . import _lib

SCHED_IDS = {
    "ddim": 0, "ddim_simple": 1, "ddim_simple_orig": 2, "ddim_simple_drag": 3, "ddpm": 4, "ddpm_orig": 5,
    "ddim_orig": 6,
}
LOGVAR_MODES = {"none": 0, "learned": 1, "fixedsmall": 2, "fixedlarge": 3}
CLIP_NONE, CLIP_CLAMP = 0, 1


class LogVar:
    """What `get_eps_logvar` hands to `pred_xprev`: the reference materialises a log-variance tensor
    (src/schedulers.py:367-390); here the formula is evaluated inside the fused update kernel, so only its
    inputs travel."""

    __slots__ = ("mode", "learned")

    def __init__(self, mode, learned):
        self.mode, self.learned = mode, learned


def _interp1d(x, y, xnew):
    """1-D linear interpolation with the reference's arithmetic (src/torchinterp1d.py:96-148): interval index
    clamp(searchsorted(x, xnew) - 1, 0, n-2), slopes (y[1:]-y[:-1]) / (eps + (x[1:]-x[:-1])) with eps =
    finfo(y.dtype).eps, result y[ind] + slopes[ind] * (xnew - x[ind]).  Returns (values, slopes)."""
    eps = torch.finfo(y.dtype).eps
    ind = torch.searchsorted(x.contiguous(), xnew.contiguous()) - 1
    ind = torch.clamp(ind, 0, x.shape[0] - 1 - 1)
    slopes = (y[1:] - y[:-1]) / (eps + (x[1:] - x[:-1]))
    return y[ind] + slopes[ind] * (xnew - x[ind]), slopes


def _even_steps(n_total, n_pick):
    """`space_timesteps(n_total, str(n_pick))` (src/schedulers.py:38-91) for a single section: n_pick indices in
    [0, n_total) at (fractional) stride (n_total-1)/(n_pick-1), rounded half-to-even like Python's round().
    A fractional n_total (continuous_t) is a section of floor(n_total)+1 steps (size_per + extra, :73-77)."""
    if n_total != int(n_total):
        n_total = n_total // 1 + 1
    if n_total < n_pick:
        raise ValueError("cannot divide section of %d steps into %d" % (n_total, n_pick))
    stride = 1 if n_pick <= 1 else (n_total - 1) / (n_pick - 1)
    out, cur = set(), 0.0
    for _ in range(n_pick):
        out.add(round(cur))
        cur += stride
    return out


def _dedup_descending(ts, max_step=999):
    """`replace_duplicate_t` (src/schedulers.py:15-31): make the timestep list strictly decreasing."""
    ts = [int(v) for v in ts]
    n = len(ts)
    up = [0] * n
    up[-2:] = ts[-2:]
    for i in range(n - 1, 0, -1):
        up[i - 1] = ts[i - 1] if ts[i - 1] > up[i] else up[i] + 1
    out = [0] * n
    ceil_t = max_step
    for i in range(n - 1):
        out[i] = min(up[i], ceil_t)
        ceil_t = out[i] - 1
    return out


class Scheduler:
    """src/schedulers.py:95-422.  Tensors live on the CPU until `.to(device)`; `sigmas` (the 1000-entry table)
    is what the device-side searchsorted reads."""

    kind = None

    def __init__(self, num_train_timesteps=1000, beta_start=0.0001, beta_end=0.02, beta_schedule="linear",
                 set_alpha_to_one=True, sampler_var="none", eta=0.0):
        n = num_train_timesteps
        if beta_schedule == "linear":
            betas = torch.linspace(beta_start, beta_end, n, dtype=torch.float32)
        elif beta_schedule == "quadratic":
            betas = torch.linspace(beta_start ** 0.5, beta_end ** 0.5, n, dtype=torch.float32) ** 2
        elif beta_schedule == "cosine":
            s = 0.008
            x = torch.linspace(0, n, n + 1)
            ac = torch.cos(((x / n) + s) / (1 + s) * torch.pi * 0.5) ** 2
            ac = ac / ac[0]
            betas = torch.clip(1 - (ac[1:] / ac[:-1]), 1e-6, 0.999)
        elif beta_schedule == "sigmoid":
            betas = torch.sigmoid(torch.linspace(-6, 6, n)) * (beta_end - beta_start) + beta_start
        else:
            raise NotImplementedError("%s is not implemented for %s" % (beta_schedule, self.__class__))
        self.betas = betas
        self.set_alpha_to_one = set_alpha_to_one
        self.num_train_timesteps = n
        self.alphas = 1.0 - betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self.final_alpha_cumprod = torch.tensor(1.0)
        self.sigmas = (1 / self.alphas_cumprod - 1).sqrt()
        self.final_sigma = (1 / self.final_alpha_cumprod - 1).sqrt()
        self.train_timesteps = torch.arange(0, n, dtype=torch.int64)
        self.timesteps = self.train_timesteps
        self.sampling_sigmas = self.sigmas
        self.continuous_t = False
        self.sampler_var = sampler_var
        self.eta = eta
        prev = torch.cat([self.final_alpha_cumprod.view(1), self.alphas_cumprod[:-1]])
        self.posterior_variance = betas * (1.0 - prev) / (1.0 - self.alphas_cumprod)
        self.min_var_coef = self.posterior_variance[1]
        self.device = torch.device("cpu")
        self.reset_state()

    # ---------------------------------------------------------------- bookkeeping
    def to(self, device):
        self.device = torch.device(device)
        for name in ("betas", "alphas_cumprod", "final_alpha_cumprod", "sigmas", "final_sigma", "train_timesteps",
                     "sampling_sigmas", "posterior_variance", "min_var_coef"):
            setattr(self, name, getattr(self, name).to(device))
        self.timesteps_host = self.timesteps.cpu()
        self.timesteps = self.timesteps.to(device)
        # continuous_t keeps numpy-born float64 sampling sigmas in the reference (src/schedulers.py:254-271); every
        # use multiplies them into float32 tensors, i.e. rounds them to float32 first: done once here
        self.sampling_sigmas = self.sampling_sigmas.to(torch.float32).contiguous()
        self.sigma_table = self.sigmas.to(torch.float32).contiguous()
        self.slopes_table = None
        if self.continuous_t:  # Interp1d(sigma table -> train timestep) slopes for the device-side time lookup
            _, sl = _interp1d(self.sigmas.cpu(), self.train_timesteps.float().cpu(), self.sigmas.cpu()[:1])
            self.slopes_table = sl.to(device, torch.float32).contiguous()
        self.min_var_coef_host = float(self.min_var_coef)
        return self

    def reset_state(self):
        self.state = {}
        self.i = 0

    # ---------------------------------------------------------------- sigma <-> t on the host (set-up only)
    def sigma_to_t(self, sigma):
        return torch.searchsorted(self.sigmas, torch.as_tensor(sigma, dtype=self.sigmas.dtype, device=self.sigmas.device))

    def sigma_to_t_interp(self, sigma):
        """src/schedulers.py:210-220 (set-up only; the per-step lookup is done by the kernels)."""
        xnew = torch.as_tensor(sigma).to(self.sigmas.device).squeeze()
        if xnew.dim() == 0:
            xnew = xnew.unsqueeze(0)
        t, _ = _interp1d(self.sigmas, self.train_timesteps.float(), xnew)
        return t.float()

    def t_to_sigma_interp(self, t):
        """src/schedulers.py:192-203."""
        xnew = t.to(self.sigmas.device).squeeze()
        if xnew.dim() == 0:
            xnew = xnew.unsqueeze(0)
        y_new, _ = _interp1d(self.train_timesteps.float(), self.alphas_cumprod, xnew)
        sigma = (1 / y_new - 1).sqrt()
        return torch.where(t >= 0, sigma, self.final_sigma).float()

    def get_t_from_sigma(self, sigma):
        if self.continuous_t:
            return self.sigma_to_t_interp(sigma)
        return self.sigma_to_t(sigma)

    def sigma(self, timestep):
        sig = self.sigmas[timestep]
        return torch.where(timestep >= 0, sig, self.final_sigma)

    def get_sigma(self, timestep):
        if self.continuous_t:
            return self.t_to_sigma_interp(timestep)
        return self.sigma(timestep)

    def set_timesteps_sigma(self, start, end, num_inference_steps, style="DDIM", scale=1, continuous_t=False):
        """src/schedulers.py:227-284."""
        self.continuous_t = bool(continuous_t)
        self.num_inference_steps = num_inference_steps
        n = num_inference_steps if self.set_alpha_to_one else num_inference_steps + 1
        sig = None
        if style == "DDIM":
            t_hi = self.get_t_from_sigma(start).item()
            t_lo = self.get_t_from_sigma(end).item()
            picks = _even_steps(t_hi + 1 - t_lo, n)
            ts = torch.tensor(t_lo + np.array(sorted(picks, reverse=True)),
                              dtype=torch.float32 if self.continuous_t else torch.long)
            if self.continuous_t:
                sig = self.get_sigma(ts)
        elif style == "EDM":
            rho = 7  # fp32 tensor arithmetic, like the reference (start/end are 0-d fp32 tensors there)
            s0, s1 = torch.as_tensor(start, dtype=torch.float32), torch.as_tensor(end, dtype=torch.float32)
            sig = torch.stack([(s0 ** (1 / rho) + i / (n - 1) * (s1 ** (1 / rho) - s0 ** (1 / rho))) ** rho
                               for i in range(n)])
            ts = self.get_t_from_sigma(sig)
        elif style == "Linear":
            # literally the reference's expression: np.log of a 0-d fp32 torch tensor is torch's fp32 log
            sig = torch.tensor(np.exp(np.linspace(np.log(start), np.log(end), n)))
            ts = self.get_t_from_sigma(sig if self.continuous_t else sig.to(self.sigmas.dtype))
        elif style == "Scaled":
            l0, l1 = np.log(start), np.log(end)
            diff = l1 - l0
            a_t = scale ** np.arange(n - 1)
            csum = np.cumsum(a_t)
            logs = np.insert(l0 + diff / csum[-1] * csum, 0, l0)
            sig = torch.tensor(np.exp(logs))
            ts = self.get_t_from_sigma(sig if self.continuous_t else sig.to(self.sigmas.dtype))
        else:
            raise ValueError("Invalid style!")
        if self.continuous_t:
            self.timesteps = ts.squeeze()
            self.sampling_sigmas = sig.squeeze()
        else:
            ts = torch.tensor(_dedup_descending(ts.squeeze().tolist()), dtype=torch.long)
            self.timesteps = ts
            self.sampling_sigmas = self.get_sigma(ts)
        if self.set_alpha_to_one:
            self.timesteps = torch.cat([self.timesteps, torch.tensor([-1])])
            self.sampling_sigmas = torch.cat([self.sampling_sigmas, torch.tensor([self.final_sigma])])
        s_t, s_p = self.sampling_sigmas[-3], self.sampling_sigmas[-2]
        beta_t = (s_t ** 2 - s_p ** 2) / (s_t ** 2 + 1)
        a_t, a_p = 1 / (s_t ** 2 + 1), 1 / (s_p ** 2 + 1)
        self.min_var_coef = beta_t * (1 - a_p) / (1 - a_t)
        if self.device.type != "cpu":
            # a schedule set after .to(device): refresh the host / device snapshots the loops and pred_xprev read
            # (timesteps_host, sigma_table, slopes_table, min_var_coef_host), which .to() takes
            self.to(self.device)

    # ---------------------------------------------------------------- per-step device work
    def get_eps_logvar(self, sigma_t, sigma_prev, learned_logvar=None):
        """Returns a LogVar handle (or None, like the reference, when sampler_var == 'none')."""
        mode = LOGVAR_MODES.get(self.sampler_var, 0)
        if mode == 0:
            return None
        if mode == 1 and learned_logvar is None:
            raise ValueError("sampler_var='learned' needs the model's variance channels")
        return LogVar(mode, learned_logvar if mode == 1 else None)

    def pred_xstart(self, xt, eps, sigma_t, clip=CLIP_NONE, out=None):
        """x0 = x_t - sigma_t * eps (src/schedulers.py:407-409), optionally fused with the clamp clip."""
        out = torch.empty_like(xt) if out is None else out
        ops.pred_xstart(xt, eps, _as_f32(sigma_t, xt.device), clip, out)
        return out

    def pred_xprev(self, x0, eps, sigma_t, sigma_prev, xt=None, log_variance=None, noise=None, out=None,
                   nan_flag=None):
        """x_{t-1} (src/schedulers.py:432-627).  `noise` defaults to the reference's own draw,
        torch.randn_like(x0) on x0's device generator, so seeds reproduce the reference stream."""
        eta = float(self.eta)
        needs_noise = self.kind in ("ddpm", "ddpm_orig") or eta > 0
        if needs_noise and noise is None:
            noise = torch.randn_like(x0)
        mode = log_variance.mode if log_variance is not None else 0
        learned = log_variance.learned if log_variance is not None else None
        out = torch.empty_like(x0) if out is None else out
        ops.pred_xprev(SCHED_IDS[self.kind], eta, x0, eps, xt, noise if needs_noise else None, learned, mode,
                       self.min_var_coef_host if hasattr(self, "min_var_coef_host") else float(self.min_var_coef),
                       _as_f32(sigma_t, x0.device), _as_f32(sigma_prev, x0.device), out, nan_flag)
        self.i += 1
        return out


def _as_f32(v, device):
    if not torch.is_tensor(v):
        v = torch.tensor(float(v))
    return v.to(device=device, dtype=torch.float32).reshape(-1).contiguous()


def _make(kind_name, force_eta=None):
    class _S(Scheduler):
        kind = kind_name

        def __init__(self, *a, **k):
            if force_eta is not None:
                k["eta"] = force_eta
            super().__init__(*a, **k)

    return _S


DDIM_Scheduler = _make("ddim")
DDIM_simple_Scheduler = _make("ddim_simple")
DDIM_simple_orig_Scheduler = _make("ddim_simple_orig")
DDIM_simple_drag_Scheduler = _make("ddim_simple_drag")
DDPM_Scheduler = _make("ddpm")
DDPM_orig_Scheduler = _make("ddpm_orig", force_eta=1.0)
DDIM_orig_Scheduler = _make("ddim_orig")

_BY_NAME = {"ddpm": DDPM_Scheduler, "ddim": DDIM_Scheduler, "ddim_simple": DDIM_simple_Scheduler,
            "ddim_orig": DDIM_orig_Scheduler, "ddim_simple_orig": DDIM_simple_orig_Scheduler,
            "ddim_simple_drag": DDIM_simple_drag_Scheduler, "ddpm_orig": DDPM_orig_Scheduler}


def get_sampler(sampler_name, train_timesteps, inference_timesteps, beta_start=0.0001, beta_end=0.02,
                beta_schedule="linear", sigma_style="DDIM", set_alpha_to_one=True, start_sigma=None, end_sigma=None,
                sampler_var="none", continuous_t=False, linear_scale=1.0, eta=0.0, ge_gamma=2, norm_eps=False,
                start_t=None, end_t=None):
    """src/schedulers.py:676-726.  'ge' (GE_Scheduler) is rejected: its pred_xprev cannot be called by the
    reference's own loops (missing `xt=` parameter)."""
    if sampler_name not in _BY_NAME:
        raise NotImplementedError(sampler_name)
    s = _BY_NAME[sampler_name](num_train_timesteps=train_timesteps, beta_start=beta_start, beta_end=beta_end,
                               beta_schedule=beta_schedule, set_alpha_to_one=set_alpha_to_one,
                               sampler_var=sampler_var, eta=eta)
    if start_sigma is None or start_sigma <= 0:
        if start_t is None or start_t < 0:
            start_sigma = s.sigmas[-1]
        else:
            start_sigma = min(s.sigmas[start_t], s.sigmas[-1])
    else:
        # torch.tensor(min(...)) in the reference keeps a Python int an int64 tensor, whose np.log is float64
        # (src/schedulers.py:717-718): the log-linear schedules depend on it
        v = min(start_sigma, s.sigmas[-1])
        start_sigma = v.clone() if torch.is_tensor(v) else torch.tensor(v)
    if end_sigma is None or end_sigma <= 0:
        end_sigma = s.sigmas[0] if (end_t is None or end_t < 0) else s.sigmas[end_t]
    s.set_timesteps_sigma(start=start_sigma, end=end_sigma, num_inference_steps=inference_timesteps,
                          style=sigma_style, scale=linear_scale, continuous_t=continuous_t)
    return s
