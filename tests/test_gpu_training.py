"""SURVEY 8f rank 3, first slice: batch preparation and the fused AdamW + EMA update of the sigma-model training step
(src/experiments.py:654-694) against the oracle (torch on the CPU: the reference's own formulas and torch.optim.AdamW)."""
import pytest
import torch

from oracle import training as OT
from oracle import weights

pytestmark = pytest.mark.gpu
dev = torch.device("cuda:0")


@pytest.mark.parametrize("shape", [(5, 3, 16, 16), (3, 3, 64, 64), (4, 7)])
def test_prepare_batch(shape):
    from nlc_b200 import training as T
    from nlc_b200.schedulers import get_sampler
    g = torch.Generator().manual_seed(3)
    B = shape[0]
    x0 = torch.rand(shape, generator=g) * 2 - 1
    noise, extra = torch.randn(shape, generator=g), torch.randn(shape, generator=g)
    sh = (B,) + (1,) * (len(shape) - 1)
    eta1, eta2 = 0.05 + torch.rand(sh, generator=g) * 0.2, 0.1 + torch.rand(sh, generator=g)
    t = torch.randint(0, 1000, (B,), generator=g)
    sch = get_sampler("ddim", 1000, 10)
    ab = sch.alphas_cumprod.cpu()
    want_x, want_d, want_n = OT.prepare_batch(x0, t, noise, extra, eta1, eta2, ab)
    got_x, got_d, got_n = T.prepare_batch(x0.to(dev), t.to(dev), noise.to(dev), extra.to(dev), eta1.to(dev), eta2.to(dev),
                                          sch.alphas_cumprod, return_noise=True)
    assert got_d.shape == want_d.shape
    assert (got_n.cpu() - want_n).abs().max() <= 1e-6 * want_n.abs().max()
    assert (got_x.cpu() - want_x).abs().max() <= 1e-6 * want_x.abs().max()
    assert ((got_d.cpu() - want_d).abs() / want_d).max() < 2e-6
    # the EDM experiment's variant (src/experiments.py:996-1001)
    sigma = (torch.randn(sh, generator=g) * 1.2 - 1.2).exp()
    want_x, want_d, want_n = OT.prepare_batch_edm(x0, sigma, noise, extra, eta1, eta2)
    got_x, got_d, got_n = T.prepare_batch_edm(x0.to(dev), sigma.to(dev), noise.to(dev), extra.to(dev), eta1.to(dev),
                                              eta2.to(dev), return_noise=True)
    assert (got_n.cpu() - want_n).abs().max() <= 1e-6 * want_n.abs().max()
    assert (got_x.cpu() - want_x).abs().max() <= 1e-6 * want_x.abs().max()
    assert ((got_d.cpu() - want_d).abs() / want_d).max() < 2e-6


@pytest.mark.parametrize("n", [4096, 10007])
def test_adamw_ema_kernel_follows_torch(n):
    import ctypes as C
    from nlc_b200 import _lib
    g = torch.Generator().manual_seed(n)
    p0 = torch.randn(n, generator=g)
    ref_p = torch.nn.Parameter(p0.clone())
    ref = OT.AdamWEma([ref_p], lr=2e-3, weight_decay=0.01, ema_rate=0.99)
    p, m, v, ema = p0.clone().to(dev), torch.zeros(n, device=dev), torch.zeros(n, device=dev), p0.clone().to(dev)
    L, st = _lib.lib(), C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for step in range(1, 7):
        grad = torch.randn(n, generator=g) * (0.1 if step % 2 else 10.0)
        ref_p.grad = grad.clone()
        ref.step()
        gd = (grad * 2).to(dev)  # the kernel folds the data-parallel mean: two ranks' summed gradients, scale 1/2
        _lib.check(L.nlc_adamw_ema_step(_lib.ctx(0), p.data_ptr(), gd.data_ptr(), m.data_ptr(), v.data_ptr(), ema.data_ptr(),
                                        n, 2e-3, 0.9, 0.999, 1e-8, 0.01, step, 0.99, 0.5, st))
        assert (p.cpu() - ref_p.detach()).abs().max() <= 2e-6 * ref_p.detach().abs().max(), step
        assert (ema.cpu() - ref.ema[0]).abs().max() <= 2e-6 * ref.ema[0].abs().max(), step
    state = ref.optim.state[ref_p]
    assert (m.cpu() - state["exp_avg"]).abs().max() <= 2e-6 * state["exp_avg"].abs().max()
    assert (v.cpu() - state["exp_avg_sq"]).abs().max() <= 2e-6 * state["exp_avg_sq"].abs().max()


class _Head(torch.nn.Module):
    """A stand-in sigma-model (the real one is the reference's nn.Module: its forward / backward is not part of this slice)."""

    def __init__(self, c):
        super().__init__()
        self.conv = torch.nn.Conv2d(c, 8, 3, padding=1)
        self.fc = torch.nn.Linear(8, 1)

    def forward(self, f):
        return self.fc(torch.nn.functional.silu(self.conv(f)).mean(dim=(2, 3))).view(-1, 1, 1, 1)


def test_train_step_end_to_end():
    """Frozen CUDA UNet encode + torch sigma-head + fused optimizer: the parameters after three iterations equal those of the
    same head trained with torch.optim.AdamW on the same features, and the EMA follows `update_ema`."""
    from nlc_b200 import training as T
    from nlc_b200.schedulers import get_sampler
    from nlc_b200.unet_ddim import UNetModel
    cfg = weights.CONFIGS["tiny"]
    model = UNetModel(**cfg["unet"], precision="tf32", device=dev).load_state_dict(
        weights.ddim_unet_state_dict(**cfg["unet"], seed=3))
    sch = get_sampler("ddim", 1000, 10).to(dev)
    torch.manual_seed(0)
    head, twin = _Head(256).to(dev), _Head(256).to(dev)
    twin.load_state_dict(head.state_dict())
    trainer = T.SigmaTrainer(head, lr=1e-3, weight_decay=0.01, ema_rate=0.9)
    ref = OT.AdamWEma(twin.parameters(), lr=1e-3, weight_decay=0.01, ema_rate=0.9)
    loss_fn = torch.nn.MSELoss()
    g = torch.Generator(device=dev).manual_seed(1)
    B, R = 6, cfg["unet"]["image_size"]
    for it in range(3):
        x = torch.rand(B, 3, R, R, generator=g, device=dev) * 2 - 1
        t = torch.randint(0, 1000, (B,), generator=g, device=dev)
        noise, extra = torch.randn(x.shape, generator=g, device=dev), torch.randn(x.shape, generator=g, device=dev)
        eta1 = 0.1 + torch.rand(B, 1, 1, 1, generator=g, device=dev) * 0.1
        eta2 = 0.5 + torch.rand(B, 1, 1, 1, generator=g, device=dev)
        loss = T.train_step(model, trainer, sch, x, t, noise, extra, eta1, eta2, loss_fn, microbatch=4)
        # the same iteration with plain torch on the twin
        noisy_x, dist_real = T.prepare_batch(x, t, noise, extra, eta1, eta2, sch.alphas_cumprod)
        with torch.no_grad():
            feat = model.encode(noisy_x, t).clone()
        ref.optim.zero_grad()
        ref_loss = loss_fn(dist_real, twin(feat) + 1)
        ref_loss.backward()
        ref.step()
        assert torch.isfinite(loss) and abs(loss.item() - ref_loss.item()) <= 1e-5 * abs(ref_loss.item())
        for (name, a), b in zip(head.named_parameters(), twin.parameters()):
            assert (a - b).abs().max() <= 1e-5 * b.abs().max().clamp_min(1e-3), (it, name)
    ema = trainer.ema_state_dict()
    for (name, _), e in zip(twin.named_parameters(), ref.ema):
        assert (ema[name] - e).abs().max() <= 1e-5 * e.abs().max().clamp_min(1e-3), name
