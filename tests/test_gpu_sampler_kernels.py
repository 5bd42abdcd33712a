"""GPU parity of the sampler arithmetic (rows D1, S2-S5) against the CPU oracle on the same seeded inputs.
fp32 elementwise work: tolerance 2e-6 relative (different FMA contraction / libm), searchsorted bit-exact."""
import math

import pytest
import torch

from oracle import sampler as S

pytestmark = pytest.mark.gpu
dev = torch.device("cuda:0")


def _rel(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-30)).item()


def test_row_norm_and_normalize():
    from nlc_b200 import ops
    g = torch.Generator().manual_seed(0)
    for shape in ((5, 3, 16, 16), (3, 3, 256, 256)):
        x = torch.randn(shape, generator=g) * 3
        n = torch.zeros(shape[0], device=dev)
        ops.row_norm(x.to(dev), n)
        assert _rel(n.cpu(), S.vector_norm(x).reshape(-1)) < 3e-6
        y = x.to(dev).clone()
        ops.normalize_rows_(y)
        assert _rel(y.cpu(), S.normalize(x, x[0].numel())) < 3e-6


def test_refine_and_correct_match_the_oracle_bit_for_bit_in_t():
    from nlc_b200 import ops
    tab = S.Tables()
    table = tab.sigmas.to(dev)
    g = torch.Generator().manual_seed(1)
    B, d = 64, 3 * 32 * 32
    x = torch.randn(B, 3, 32, 32, generator=g) * torch.linspace(0.05, 120, B).view(B, 1, 1, 1)
    norms = torch.zeros(B, device=dev)
    ops.row_norm(x.to(dev), norms)
    for sigma0 in (0.5, 7.0, 80.0):
        nmin, nmax = -2.0 / math.sqrt(d), 0.9
        nx = S.vector_norm(x) / math.sqrt(d)
        ref_sigma = torch.clamp(torch.ones_like(nx) * sigma0, min=torch.clamp(nx - nmax, min=0), max=nx + nmin)
        ref_t = torch.clamp(tab.t_of_sigma(ref_sigma), min=0.0, max=1000.0).reshape(-1)
        sig, t, sc = (torch.zeros(B, device=dev) for _ in range(3))
        ops.refine_sigma(norms, B, d, torch.tensor([sigma0], device=dev), nmin, nmax, True, 0.0, table, 0, sig, t, sc)
        # sigma goes through the same fp32 ops; the norm feeding the clamp may differ in the last bit
        assert _rel(sig.cpu(), ref_sigma.reshape(-1)) < 1e-6
        exact = tab.t_of_sigma(sig.cpu()).clamp(0, 1000).float()
        assert torch.equal(t.cpu(), exact)              # searchsorted itself is bit-exact
        assert (t.cpu() - ref_t).abs().max() <= 1       # and at most one bucket from the oracle's own sigma
        r = (torch.rand(B, generator=g) - 0.5) * 0.4
        sp = torch.tensor([sigma0 * 0.9])
        sh, sph, th, sch = (torch.zeros(B, device=dev) for _ in range(4))
        ops.sigma_correct(r.to(dev), sig, sp.to(dev), True, table, sh, sph, th, sch)
        s_cpu = sig.cpu()
        dist = s_cpu * (1 + r)
        assert torch.equal(sh.cpu(), dist)
        assert torch.equal(sph.cpu(), dist * (sp / s_cpu))
        assert torch.equal(th.cpu(), tab.t_of_sigma(dist).clamp(0, 1000).float())
        assert _rel(sch.cpu(), (1 / (dist ** 2 + 1)).sqrt()) < 1e-6


KINDS = [("ddim", 0.0, "none"), ("ddim", 0.5, "fixedsmall"), ("ddim", 0.7, "learned"), ("ddim_simple", 0.2, "none"),
         ("ddim_simple_orig", 0.85, "none"), ("ddim_simple_drag", 0.2, "none"), ("ddpm", 1.0, "fixedlarge"),
         ("ddpm", 1.0, "learned"), ("ddpm_orig", 1.0, "fixedsmall"), ("ddim_orig", 0.3, "fixedlarge")]


@pytest.mark.parametrize("kind,eta,var", KINDS)
@pytest.mark.parametrize("per_sample", [False, True])
def test_pred_xstart_and_xprev(kind, eta, var, per_sample):
    from nlc_b200 import ops
    from nlc_b200.schedulers import LOGVAR_MODES, SCHED_IDS
    g = torch.Generator().manual_seed(2)
    B, shape = 4, (4, 3, 16, 16)
    xt = torch.randn(shape, generator=g) * 5
    eps = torch.randn(shape, generator=g)
    noise = torch.randn(shape, generator=g)
    v = torch.rand(shape, generator=g) * 2 - 1
    if per_sample:
        st = torch.tensor([5.0, 0.7, 30.0, 0.02]).view(B, 1, 1, 1)
        sp = torch.tensor([4.2, 0.5, 27.0, 0.0]).view(B, 1, 1, 1)
    else:
        st, sp = torch.tensor(3.0), torch.tensor(2.4)
    mvc = torch.tensor(3e-5)
    x0_ref = (xt - st * eps).clamp(-1, 1)
    x0 = torch.zeros(shape, device=dev)
    ops.pred_xstart(xt.to(dev), eps.to(dev), st.reshape(-1).to(dev), 1, x0)
    assert _rel(x0.cpu(), x0_ref) < 1e-6
    lv = S.eps_logvar(st, sp, mvc, var, v)
    need_noise = kind in ("ddpm", "ddpm_orig") or eta > 0
    ref = S.pred_xprev(kind, eta, x0_ref, eps, st, sp, xt, lv, noise if need_noise else None)
    out = torch.zeros(shape, device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    ops.pred_xprev(SCHED_IDS[kind], eta, x0_ref.to(dev), eps.to(dev), xt.to(dev), noise.to(dev) if need_noise else None,
                   v.to(dev) if var == "learned" else None, LOGVAR_MODES[var], float(mvc), st.reshape(-1).to(dev),
                   sp.reshape(-1).to(dev), out, flag)
    assert _rel(out.cpu(), ref) < 3e-6
    assert int(flag.item()) == 0


def test_nan_flag_is_raised():
    from nlc_b200 import ops
    shape = (2, 3, 8, 8)
    x0 = torch.zeros(shape, device=dev)
    x0[1, 0, 0, 0] = float("nan")
    out = torch.zeros(shape, device=dev)
    flag = torch.zeros(1, dtype=torch.int32, device=dev)
    one = torch.ones(1, device=dev)
    ops.pred_xprev(0, 0.0, x0, torch.zeros(shape, device=dev), None, None, None, 0, 0.0, one, one, out, flag)
    assert int(flag.item()) == 1
