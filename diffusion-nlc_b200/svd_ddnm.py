"""Host-side mirror of the reference's DDNM / DDNM+ samplers (functions/svd_ddnm.py) on libnlc_b200 kernels
(SURVEY §8f rank 2).

Same names, arguments and return values as the reference: `ddnm_diffusion` (:19-78), `ddnm_plus_diffusion` (:80-145),
`get_schedule_jump` (:167-190), `compute_alpha` (:10-13), `inverse_data_transform` (:15-17).  What changes underneath:
every reverse step after the network call is ONE operator call (`A_functions.ddnm_step`, C ABI `nlc_ddnm_step`): the
x0 estimate, the projection x0 - [Lambda] A^+(A x0 - y) and the x_{t-1} assembly with the (spectrally shaped) noise
are fused instead of the reference's chain of ~60 elementwise / permute / matmul launches, and the trajectory stays
on the device instead of bouncing to the CPU after every step (`xs.append(xt_next.to('cpu'))`, :65,131).  Only the
last `x_t` and `x0_t` are returned by the reference, so only those are kept.

`noise_fn(like)` (default `torch.randn_like`, i.e. the device generator exactly as in the reference) lets a test feed
recorded draws; `to_cpu=False` returns the device tensors (the reference always moves the results to the host).
"""
import ctypes as C

import torch

from . import _lib

class_num = 951  # :7


def compute_alpha(beta, t):
    """alpha_bar at t (:10-13): cumprod(1 - beta) with a leading 1, read at t + 1; shape [B,1,1,1]."""
    beta = torch.cat([torch.zeros(1).to(beta.device), beta], dim=0)
    return (1 - beta).cumprod(dim=0).index_select(0, t + 1).view(-1, 1, 1, 1)


def inverse_data_transform(x):
    return torch.clamp((x + 1.0) / 2.0, 0.0, 1.0)


def get_schedule_jump(T_sampling, travel_length, travel_repeat):
    """RePaint time-travel schedule (:167-190), including the reference's sanity checks (:192-206)."""
    remaining = {j: travel_repeat - 1 for j in range(0, T_sampling - travel_length, travel_length)}
    ts, t = [], T_sampling
    while t >= 1:
        t -= 1
        ts.append(t)
        if remaining.get(t, 0) > 0:
            remaining[t] -= 1
            for _ in range(travel_length):
                t += 1
                ts.append(t)
    ts.append(-1)
    assert ts[0] > ts[1], (ts[0], ts[1])
    for a, b in zip(ts[:-1], ts[1:]):
        assert abs(a - b) == 1, (a, b)
    assert min(ts) >= -1 and max(ts) <= T_sampling
    return ts


def _alpha_table(b):
    """compute_alpha's table, taken once per call instead of once per step, with the same torch ops on the same device
    as the reference (a GPU cumprod is a scan and may differ from the CPU's in the last bit)."""
    beta = torch.cat([torch.zeros(1).to(b.device), b], dim=0)
    return (1 - beta).cumprod(dim=0).float().cpu()


def _loop(x, model, b, eta, A_funcs, y, sigma_y, cls_fn, classes, config, noise_fn, to_cpu=True):
    skip = config.diffusion.num_diffusion_timesteps // config.time_travel.T_sampling
    times = get_schedule_jump(config.time_travel.T_sampling, config.time_travel.travel_length,
                              config.time_travel.travel_repeat)
    alphas = _alpha_table(b)
    dev = x.device
    n = x.size(0)
    xt, x0_t = x.float(), None
    y = y.reshape(n, -1).float().to(dev)
    lib = _lib.lib()
    ctx = _lib.ctx(dev.index if dev.index is not None else torch.cuda.current_device())
    with torch.no_grad():
        for i, j in zip(times[:-1], times[1:]):
            i, j = i * skip, j * skip
            if j < 0:
                j = -1
            at_next = float(alphas[j + 1])
            if j < i:  # normal sampling
                at = float(alphas[i + 1])
                t = (torch.ones(n) * i).to(dev)
                if cls_fn is None:
                    et = model(xt, t)
                else:  # classifier guidance (:46-50): host-side glue around the user's classifier gradient
                    classes = torch.ones(n, dtype=torch.long, device=dev) * class_num
                    et = model(xt, t, classes)[:, :3]
                    et = et - (1 - at) ** 0.5 * cls_fn(x, t, classes)
                x0_t, xt = A_funcs.ddnm_step(xt, et, noise_fn(xt), y, at, at_next, eta, sigma_y)
            else:  # time travel back (:67-73)
                z = noise_fn(x0_t).contiguous()
                out = torch.empty_like(x0_t)
                _lib.check(lib.nlc_ddnm_renoise(ctx, x0_t.data_ptr(), z.data_ptr(), x0_t.numel(), at_next, out.data_ptr(),
                                                C.c_void_p(torch.cuda.current_stream().cuda_stream)))
                xt = out
    if not to_cpu:  # sharded runs all-gather the device tensors
        return [xt], [x0_t]
    return [xt.to("cpu")], [x0_t.to("cpu")]


def ddnm_diffusion(x, model, b, eta, A_funcs, y, cls_fn=None, classes=None, config=None, noise_fn=torch.randn_like,
                   to_cpu=True):
    """functions/svd_ddnm.py:19-78."""
    return _loop(x, model, b, eta, A_funcs, y, None, cls_fn, classes, config, noise_fn, to_cpu)


def ddnm_plus_diffusion(x, model, b, eta, A_funcs, y, sigma_y, cls_fn=None, classes=None, config=None,
                        noise_fn=torch.randn_like, to_cpu=True):
    """functions/svd_ddnm.py:80-145."""
    return _loop(x, model, b, eta, A_funcs, y, float(sigma_y), cls_fn, classes, config, noise_fn, to_cpu)
