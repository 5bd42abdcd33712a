"""First slice of the sigma-model training step (src/experiments.py:632-700) on libnlc_b200.

One iteration of the reference: perturb the noise and diffuse the batch (:658-669), run the FROZEN UNet's `encode` under
no-grad (:673-682; 99 % of the step's FLOPs), sigma-model forward, loss against the true noise level, backward (:683-691),
AdamW on the master parameters and an EMA copy (:692-694).  Built natively here: the batch preparation (`nlc_train_prepare`,
one pass), the encoder (the sampling path's engine, any of the three network classes) and the optimizer + EMA update
(`nlc_adamw_ema_step`, one pass over a flat buffer).  NOT built: the sigma-model's own forward / backward — it stays an
ordinary `torch.nn.Module` under autograd (pass the reference's module, or any module with the same signature).

Data parallelism: the reference wraps the sigma-model in DDP but runs forward and backward under `no_sync()` (:683-687), so
its ranks never average their gradients.  `SigmaTrainer.step` all-reduces the flat gradient buffer over NCCL (one
collective per step, folded mean) before the update, which is what the DDP wrapper was meant to do.
"""
import ctypes as C

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
