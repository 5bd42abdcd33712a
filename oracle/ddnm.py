"""ORACLE (test infrastructure, not product code) — CPU fp32 restatement of the reference's DDNM / DDNM+ samplers
(functions/svd_ddnm.py): the RePaint-style time-travel schedule `get_schedule_jump` (:167-190), the noiseless loop
`ddnm_diffusion` (:19-78: x0 <- x0 - A^+(A x0 - y), DDIM step with eta) and the noisy-measurement loop
`ddnm_plus_diffusion` (:80-145: Eq. 17 / Eq. 51 with the operators' Lambda / Lambda_noise).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
Pinned by tests/golden/ddnm_plus_r32.pt (per-step dumps of the unmodified reference run on the CPU; the generator
tests/golden/make_golden.py::ddnm only redirects the loop's hard-coded `.to('cuda')` moves) and live against the reference
in tests/test_oracle_vs_reference.py.

`noise_fn(like)` replaces the reference's `torch.randn_like` so that a test can feed the same draws to this loop, to the
reference and to the CUDA path; it is called once per step, in the reference's order.
"""
import torch


def schedule_jump(T_sampling, travel_length, travel_repeat):
    """Descending step indices T-1 .. 0, -1 with `travel_repeat - 1` detours of `travel_length` steps back up after
    every `travel_length`-th index (functions/svd_ddnm.py:167-190)."""
    budget = {j: travel_repeat - 1 for j in range(0, T_sampling - travel_length, travel_length)}
    out, t = [], T_sampling
    while t >= 1:
        t -= 1
        out.append(t)
        if budget.get(t, 0) > 0:
            budget[t] -= 1
            out.extend(range(t + 1, t + travel_length + 1))
            t += travel_length
    out.append(-1)
    return out


def alpha_bar(betas, t):
    """compute_alpha (:10-13): cumulative product of (1 - beta) with a leading 1, read at t + 1."""
    padded = torch.cat([torch.zeros(1, dtype=betas.dtype), betas])
    return (1 - padded).cumprod(dim=0)[int(t) + 1]


def run(x, model, betas, eta, op, y, sigma_y=None, num_diffusion_timesteps=1000, T_sampling=100, travel_length=1,
        travel_repeat=1, noise_fn=torch.randn_like, record=None):
    """ddnm_diffusion (sigma_y is None) / ddnm_plus_diffusion.  Returns (x_last, x0_last); `record`, when a dict of
    lists, receives xt / et / x0 / x_next of every step."""
    skip = num_diffusion_timesteps // T_sampling
    times = schedule_jump(T_sampling, travel_length, travel_repeat)
    n = x.shape[0]
    xt, x0_t = x, None
    with torch.no_grad():
        for i, j in zip(times[:-1], times[1:]):
            i, j = i * skip, j * skip
            if j < 0:
                j = -1
            at_next = alpha_bar(betas, j)
            if j < i:
                at = alpha_bar(betas, i)
                et = model(xt, torch.ones(n) * i)
                if et.shape[1] == 6:
                    et = et[:, :3]
                x0_t = (xt - et * (1 - at).sqrt()) / at.sqrt()
                resid = op.A_pinv(op.A(x0_t.reshape(n, -1)) - y.reshape(n, -1))
                z = noise_fn(x0_t)
                if sigma_y is None:
                    x0_hat = x0_t - resid.reshape(x0_t.shape)
                    c1 = (1 - at_next).sqrt() * eta
                    c2 = (1 - at_next).sqrt() * ((1 - eta ** 2) ** 0.5)
                    x_next = at_next.sqrt() * x0_hat + c1 * z + c2 * et
                else:
                    sigma_t = (1 - at_next).sqrt()
                    x0_hat = x0_t - op.Lambda(resid.reshape(n, -1), at_next.sqrt(), sigma_y, sigma_t, eta).reshape(
                        x0_t.shape)
                    x_next = at_next.sqrt() * x0_hat + op.Lambda_noise(
                        z.reshape(n, -1), at_next.sqrt(), sigma_y, sigma_t, eta, et.reshape(n, -1)).reshape(x0_t.shape)
                if record is not None:
                    record["xt"].append(xt.clone()), record["et"].append(et.clone())
                    record["x0"].append(x0_t.clone()), record["x_next"].append(x_next.clone())
            else:  # time travel: re-noise the last x0 estimate up to level j
                x_next = at_next.sqrt() * x0_t + noise_fn(x0_t) * (1 - at_next).sqrt()
                if record is not None:
                    record["xt"].append(xt.clone()), record["et"].append(torch.zeros_like(xt))
                    record["x0"].append(x0_t.clone()), record["x_next"].append(x_next.clone())
            xt = x_next
    return xt, x0_t
