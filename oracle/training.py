"""ORACLE (test infrastructure, not product code) — the arithmetic of one sigma-model training iteration of the reference
(src/experiments.py:654-694) around the sigma-model's forward / backward, restated with plain torch on the CPU.

Only tests/ may import this.  The batch preparation follows :661-669 and `Scheduler.diffusion` (src/schedulers.py:323-329)
and is pinned live against the reference's scheduler in tests/test_oracle_vs_reference.py; the optimizer side IS
`torch.optim.AdamW` (:144) and `update_ema` (:233-236, src/nn_util.py:55-65), which the reference calls directly, so the
oracle calls the same torch classes.  The sigma-model's train-mode forward (batch-statistics BatchNorm) is
oracle/ddim_net.sigma_forward(training=True); tests/golden/train_step_tiny.pt pins a whole reference iteration (loss,
gradients, updated parameters, EMA) against it.
"""
import numpy as np
import torch


def prepare_batch(x0, t, noise, extra, eta1, eta2, alphas_cumprod):
    """-> (noisy_x, dist_real, new_noise): src/experiments.py:666-669."""
    noise_delta = eta1 * noise + eta1 * eta2 * extra
    new_noise = noise + noise_delta
    dim = x0[0].numel()
    dims = tuple(range(1, x0.dim()))
    dist_real = torch.linalg.vector_norm(new_noise, dim=dims, keepdim=True) / np.sqrt(dim)
    alpha = alphas_cumprod[t].view((-1,) + (1,) * (x0.dim() - 1))
    noisy_x = x0 * alpha.sqrt() + new_noise * (1 - alpha).sqrt()
    return noisy_x, dist_real, new_noise


def prepare_batch_edm(x0, sigma, noise, extra, eta1, eta2):
    """-> (noisy_img, dist_real, new_noise): src/experiments.py:998-1001."""
    noise_delta = eta1 * (noise + eta2 * extra)
    new_noise = noise + noise_delta
    dims = tuple(range(1, x0.dim()))
    dist_real = torch.linalg.vector_norm(new_noise, dim=dims, keepdim=True) / np.sqrt(x0[0].numel())
    return x0 + sigma * new_noise, dist_real, new_noise


class AdamWEma:
    """torch.optim.AdamW + the EMA copy of the parameters, as set_optimizers / update_ema use them."""

    def __init__(self, params, lr, weight_decay=0.0, betas=(0.9, 0.999), eps=1e-8, ema_rate=0.999):
        self.params = list(params)
        self.optim = torch.optim.AdamW(self.params, lr=lr, weight_decay=weight_decay, betas=betas, eps=eps)
        self.ema = [p.detach().clone() for p in self.params]
        self.rate = ema_rate

    def step(self):
        self.optim.step()
        for targ, src in zip(self.ema, self.params):
            targ.detach().mul_(self.rate).add_(src.detach(), alpha=1 - self.rate)
