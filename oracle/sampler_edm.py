"""ORACLE (test infrastructure, not product code) — CPU restatement of the reference's EDM sampling arithmetic:
`EDMImageExperiment.get_denoise_vector` / `encode_edm` / `pred_edm` (src/experiments.py:777-843) and
`edm_sampler` (:847-918), as plain functions of three callables
    net(x32, c_noise[B]) -> F,   enc(x32, c_noise[B]) -> feat,   sig(feat) -> r [B,1,1,1].
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
Pinned against the unmodified reference by tests/test_oracle_vs_reference.py::test_edm_sampler (torch.equal) and by
tests/golden/edm_sampler_tiny.pt.  Tensor dtypes follow the reference exactly: x float64, network float32, noise
levels 0-d float64 until a per-sample float32 correction makes them [B,1,1,1].
"""
import math

import numpy as np
import torch


def _single(t):
    return len(t.unsqueeze(-1)) == 1


def _norm(x):
    return torch.linalg.vector_norm(x, dim=tuple(range(1, x.dim())), keepdim=True)


def normalize(x, dim, eps=1e-12):
    return math.sqrt(dim) * x / torch.clamp(_norm(x), min=eps)


class EDM:
    def __init__(self, net, enc, sig, dim, sigma_data=0.5, norm_min=0.0, norm_max=1.0):
        self.net, self.enc, self.sig, self.dim, self.sd = net, enc, sig, dim, sigma_data
        self.norm_min, self.norm_max = norm_min, norm_max  # already divided by sqrt(dim) (:176-184)

    def encode(self, xt, sigma):  # :777-786
        xt = xt.to(torch.float32)
        sigma = sigma.to(torch.float32).reshape(-1, 1, 1, 1)
        c_in = 1 / (self.sd ** 2 + sigma ** 2).sqrt()
        c_noise = sigma.log() / 4
        return self.enc((c_in * xt).to(torch.float32), c_noise.flatten())

    def pred(self, xt, sigma):  # :788-802
        xt = xt.to(torch.float32)
        sigma = sigma.to(torch.float32).reshape(-1, 1, 1, 1)
        c_skip = self.sd ** 2 / (sigma ** 2 + self.sd ** 2)
        c_out = sigma * self.sd / (sigma ** 2 + self.sd ** 2).sqrt()
        c_in = 1 / (self.sd ** 2 + sigma ** 2).sqrt()
        c_noise = sigma.log() / 4
        F_x = self.net((c_in * xt).to(torch.float32), c_noise.flatten())
        return c_skip * xt + c_out * F_x.to(torch.float32)

    def denoise_vector(self, xt, sigma_t, sigma_prev, style="base", norm_eps=False, refine=False):  # :805-843
        orig = sigma_t
        if refine:
            nx = _norm(xt) / math.sqrt(self.dim)
            lo, hi = torch.clamp(nx - self.norm_max, min=0), nx + self.norm_min
            raw = torch.ones_like(nx) * sigma_t if _single(sigma_t) else sigma_t
            sigma_t = torch.clamp(raw, min=lo, max=hi)
            if _single(sigma_prev):
                sigma_prev = torch.ones_like(nx) * sigma_prev
        if "pred" in style:
            r = self.sig(self.encode(xt, sigma_t))
            hat = sigma_t * (1 + r)
            prev_hat = hat * (sigma_prev / sigma_t)
            sigma_t = hat
            if style == "pred":
                sigma_prev = prev_hat
        if _single(orig):
            orig = orig.reshape(-1, 1, 1, 1)
        if _single(sigma_t):
            sigma_t = sigma_t.reshape(-1, 1, 1, 1)
        if _single(sigma_prev):
            sigma_prev = sigma_prev.reshape(-1, 1, 1, 1)
        used = orig if style == "pred_sigma" else sigma_t
        denoised = self.pred(xt, used).to(torch.float64)
        eps = (xt - denoised) / used
        if norm_eps:
            eps = normalize(eps, self.dim)
        return eps, denoised, sigma_t, sigma_prev

    def sigma_steps(self, num_steps, sigma_min, sigma_max, rho=7, scheduler="EDM"):  # :860-868
        idx = torch.arange(num_steps, dtype=torch.float64)
        if scheduler == "EDM":
            s = (sigma_max ** (1 / rho) + idx / (num_steps - 1) * (sigma_min ** (1 / rho) - sigma_max ** (1 / rho))) ** rho
        else:
            s = torch.tensor(np.exp(np.linspace(np.log(sigma_max), np.log(sigma_min), num_steps)))
        return torch.cat([torch.as_tensor(s), torch.zeros_like(s[:1])])

    def sample(self, latents, num_steps, sigma_min=0.002, sigma_max=80, rho=7, style="base,base", norm_eps="000",
               refine=False, scheduler="EDM", eps_ratio=0.5, eps_scale=1.0, second_order=True, log=None):  # :847-918
        ne, nec = bool(int(norm_eps[0])), bool(int(norm_eps[1]))
        st, sn = style.split(",")
        steps = self.sigma_steps(num_steps, sigma_min, sigma_max, rho, scheduler)
        cos = torch.nn.CosineSimilarity(dim=1, eps=1e-6)
        x_next = latents.to(torch.float64) * steps[0]
        for i, (cur, nxt) in enumerate(zip(steps[:-1], steps[1:])):
            x_hat = x_next  # S_churn = 0: gamma = 0, the churn term is exactly zero (:877-880)
            nxt0 = nxt
            hat0 = torch.as_tensor(cur + 0 * cur)
            eps, _, hat, nxt = self.denoise_vector(x_hat, hat0, nxt, st, ne, refine)
            eps = eps * (hat / hat0)
            if "pred_partial" in st:
                nxt = nxt0
            x_next = x_hat + ((nxt - hat0) if st == "pred_partial" else (nxt - hat)) * eps
            if st == "pred_partial3":
                hat = hat0
            if i < num_steps - 1 and second_order:
                eps2, _, nxt, _ = self.denoise_vector(x_next, nxt, nxt * 0, sn, ne, refine)
                eps2 = eps2 * (nxt / nxt0)
                if "pred_partial" in sn:
                    nxt = nxt0
                new = eps_ratio * eps + (1 - eps_ratio) * eps2
                if nec:
                    new = normalize(new, self.dim)
                if eps_scale is not None:
                    new = new / eps_scale
                else:
                    b = len(new)
                    new = new * cos(new.reshape(b, -1), eps.reshape(b, -1)).reshape(b, 1, 1, 1)
                x_next = x_hat + (nxt - hat) * new
            if log is not None:
                log.append(dict(x_hat=x_hat.clone(), x_next=x_next.clone(),
                                sigma_hat=torch.as_tensor(hat).reshape(-1).clone().double(),
                                eps=eps.clone()))
        return x_next
