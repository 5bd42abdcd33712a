"""Host-side logic that needs no GPU: scheduler mirror vs oracle / golden tables, weight packing, activation
views, and the world_size-2 sharding path over gloo."""
import os

import pytest
import torch
import torch.multiprocessing as mp

from oracle import sampler as S


def test_scheduler_mirror_matches_golden(golden_dir):
    from nlc_b200 import schedulers as M
    g = torch.load(os.path.join(golden_dir, "scheduler_tables.pt"), weights_only=True)
    for name, kw in (("ddim50_s100", dict(sampler_name="ddim", inference_timesteps=50, start_sigma=100)),
                     ("simple_orig100", dict(sampler_name="ddim_simple_orig", inference_timesteps=100,
                                             start_sigma=100, eta=0.85)),
                     ("ddim6_s20", dict(sampler_name="ddim", inference_timesteps=6, start_sigma=20.0))):
        s = M.get_sampler(train_timesteps=1000, **kw)
        assert torch.equal(s.timesteps, g[name]["timesteps"])
        assert torch.equal(s.sampling_sigmas.float(), g[name]["sigmas"])
        assert torch.equal(s.sigmas, g[name]["table"])
        assert float(s.min_var_coef) == float(g[name]["min_var_coef"])


def test_scheduler_mirror_matches_oracle_tables():
    from nlc_b200 import schedulers as M
    tab = S.Tables()
    for n, start in ((10, 50.0), (100, 100.0), (37, 3.0)):
        s = M.get_sampler("ddim", 1000, n, start_sigma=start)
        ts, sig, mvc = tab.ddim_schedule(start, None, n)
        assert torch.equal(s.timesteps, ts) and torch.equal(s.sampling_sigmas, sig)
        assert float(s.min_var_coef) == float(mvc)


def test_unknown_sampler_and_ge_are_rejected():
    from nlc_b200 import schedulers as M
    with pytest.raises(NotImplementedError):
        M.get_sampler("ge", 1000, 10)
    with pytest.raises(NotImplementedError):
        M.get_sampler("nope", 1000, 10)


def test_pack_conv_weight_order():
    from nlc_b200 import ops
    from nlc_b200._lib import NLC_BF16, NLC_F32
    w = torch.arange(2 * 3 * 3 * 3, dtype=torch.float32).reshape(2, 3, 3, 3)
    k = ops.pack_conv_weight(w, NLC_F32)
    assert k.shape == (2, 27)
    # K index = (kh*3 + kw)*Cin + c
    assert k[1, (1 * 3 + 2) * 3 + 1] == w[1, 1, 1, 2]
    extra = torch.ones(2, 5, 1, 1)
    assert ops.pack_conv_weight(w, NLC_BF16, extra=extra).shape == (2, 32)


def test_tf32_rounding_matches_definition():
    from nlc_b200 import ops
    x = torch.tensor([1.0 + 2 ** -11, 1.0 + 2 ** -12, -3.1415927, 1e-30, 65504.0])
    r = ops.round_tf32_(x.clone())
    assert (r.view(torch.int32) & 0x1FFF).abs().sum() == 0
    assert ((r - x).abs() <= x.abs() * 2 ** -11).all()
    assert r[0] == 1.0 + 2 ** -10  # tie rounds away from zero


def test_act_view_slicing():
    from nlc_b200.ops import Act
    t = torch.zeros(2, 4, 4, 96)
    a = Act(t, 32, 64)
    assert (a.B, a.H, a.W, a.C, a.ld) == (2, 4, 4, 64, 96)
    assert a.ptr == t.data_ptr() + 32 * 4
    assert a.slice(16, 8).ptr == t.data_ptr() + 48 * 4


def test_shard_ranges_cover_the_batch():
    from nlc_b200.parallel import shard_range
    for B in (1, 7, 256, 257):
        for ws in (1, 2, 3, 8):
            spans = [shard_range(B, r, ws) for r in range(ws)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(ws - 1))


def _worker(rank, ws, port, q):
    import torch.distributed as dist
    from nlc_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    B, shape = 5, (5, 3, 4, 4)
    full = torch.randn(shape, generator=torch.Generator().manual_seed(77))
    mine, = parallel.sharded_noise(shape, 77, rank, ws)
    lo, hi = parallel.shard_range(B, rank, ws)
    ok = torch.equal(mine, full[lo:hi])
    # "finished images": every rank doubles its rows; the gather must equal the un-sharded result
    gathered = parallel.gather_images(mine * 2, global_batch=B)
    ok = ok and torch.equal(gathered, full * 2)
    loss, nan, tmax = parallel.global_step_scalars(mine.abs().sum(), float(rank == 1), float(10 + rank))
    ok = ok and abs(float(loss) - float(full.abs().sum())) < 1e-3 and bool(nan) and float(tmax) == 10 + ws - 1
    sums = parallel.reduce_metric_sums(torch.tensor([float(hi - lo), 1.0]))
    ok = ok and sums.tolist() == [float(B), float(ws)]
    # evaluation metrics of a sharded run (image_sample.evaluate_constraint): global means over uneven shards, incl. a rank
    # whose shard is empty for one key set
    from nlc_b200 import metrics
    vals = {"mse": torch.arange(lo, hi, dtype=torch.float64), "ssim": torch.arange(lo, hi, dtype=torch.float64) * 2,
            "psnr": torch.ones(hi - lo, dtype=torch.float64)}
    means = metrics.reduce_means(vals, keys=("mse", "psnr", "ssim"))
    ok = ok and abs(means["mse"] - (B - 1) / 2) < 1e-12 and abs(means["ssim"] - (B - 1)) < 1e-12 and means["psnr"] == 1.0
    # round-robin dealing of batches to ranks, as the evaluation drivers do it
    dealt = [i for i in range(7) if i % ws == rank]
    counts = parallel.reduce_metric_sums(torch.tensor([float(len(dealt))]))
    ok = ok and counts.item() == 7.0
    # FID statistics of a sharded run (fid.FidStatistics.all_reduce / finalize): each rank holds the fp64 partial sums of its
    # own features (filled directly here: the accumulation kernel needs the GPU); the reduced result is np.mean / np.cov of
    # the union
    from nlc_b200 import fid
    import numpy as np
    feats = np.random.default_rng(11).normal(size=(10, 6))
    mine = torch.from_numpy(feats[rank::ws])
    st = fid.FidStatistics(dims=6, device="cpu")
    st.sum += mine.sum(0)
    st.outer += mine.T @ mine
    st.count = mine.shape[0]
    mu, sigma = st.all_reduce().finalize()
    ok = ok and st.count == 10 and np.allclose(mu, feats.mean(0)) and np.allclose(sigma, np.cov(feats, rowvar=False))
    q.put((rank, ok))
    dist.destroy_process_group()


def test_world_size_2_sharding_over_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_upsample_phase_weights_reproduce_upsample_then_conv():
    """ops.upsample_phase_weights / upsample_phase_taps: the four sub-pixel phase kernels (summed rows / columns of a 3x3
    kernel over two source rows / columns) are "nearest x2, then 3x3 conv with padding 1" (src/unet_ddim.py:58-74), checked
    against torch on the CPU with the packed weights exactly as the kernels receive them (fp32 container)."""
    import torch.nn.functional as F
    from nlc_b200 import ops
    from nlc_b200._lib import NLC_F32X3
    g = torch.Generator().manual_seed(5)
    B, Cin, Cout, H, W = 2, 8, 12, 6, 10
    w, x = torch.randn(Cout, Cin, 3, 3, generator=g), torch.randn(B, Cin, H, W, generator=g)
    want = F.conv2d(F.interpolate(x, scale_factor=2, mode="nearest"), w, padding=1)
    got = torch.zeros_like(want)
    packs = ops.upsample_phase_weights(w, NLC_F32X3)  # plain fp32 packing: [Cout, tap, Cin]
    xp = F.pad(x, (1, 1, 1, 1))
    for (a, b), pk in packs.items():
        k = pk.reshape(Cout, 4, Cin)
        taps = ops.upsample_phase_taps(0, 0, Cin, a, b)
        assert [t[3:] for t in taps] == [(0, Cin)] * 4 and len(taps) == 4
        acc = 0
        for i, (_, dh, dw, _, _) in enumerate(taps):
            acc = acc + torch.einsum("oc,nchw->nohw", k[:, i], xp[:, :, 1 + dh:1 + dh + H, 1 + dw:1 + dw + W])
        got[:, :, a::2, b::2] = acc
    assert (got - want).abs().max() < 1e-5 * want.abs().max()


def test_ddnm_schedule_and_factory_helpers_without_a_gpu():
    """Host logic of the SURVEY 8f modules that needs no device: the RePaint time-travel schedule equals the oracle's (and the
    reference's own checks hold), channel-multiplier defaults of the factories (src/script_util.py:158-172), file-resume
    helper of the evaluation drivers."""
    from nlc_b200 import image_sample as IS, script_util as SU, svd_ddnm as SD
    from oracle import ddnm as OD
    for T, length, repeat in ((4, 2, 2), (10, 3, 3), (100, 1, 1), (25, 10, 2), (20, 5, 1)):
        ts = SD.get_schedule_jump(T, length, repeat)
        assert ts == OD.schedule_jump(T, length, repeat)
        assert ts[0] == T - 1 and ts[-1] == -1 and all(abs(a - b) == 1 for a, b in zip(ts[:-1], ts[1:]))
        # every index with budget is revisited (repeat - 1) more times
        assert ts.count(0) == (repeat if length < T else 1)
    assert SU._channel_mult("", 256) == (1, 1, 2, 2, 4, 4) and SU._channel_mult("", 32) == (1, 2, 2, 2)
    assert SU._channel_mult("1,2,3", 64) == (1, 2, 3) and SU._channel_mult((1, 2), 64) == (1, 2)
    with pytest.raises(ValueError):
        SU._channel_mult("", 48)
    assert IS._already_done(None, 0, 0, 4) is False
    a = SD.compute_alpha(torch.linspace(1e-4, 2e-2, 1000), torch.tensor([0, 499, -1]))
    assert a.shape == (3, 1, 1, 1) and float(a[2]) == 1.0 and float(a[0]) == float(1 - torch.tensor(1e-4))


def test_batch_streams_shards_reproduce_the_unsharded_noise():
    """image_sample.BatchStreams: with batches dealt round-robin to 2 ranks, each rank replays the global x_T stream (CPU
    generator) and the per-step device stream and discards the draws of the batches it does not own, so the union of the
    shards equals the un-sharded run (and no two ranks start from the same noise)."""
    import types
    from nlc_b200 import image_sample as IS
    shape, n_batches, n_steps = (3, 2, 4, 4), 5, 3

    def run(rank, world):
        exp = types.SimpleNamespace(device=torch.device("cpu"), new_gen=lambda: torch.manual_seed(11))
        st = IS.BatchStreams(exp, shape, rank, world, device_draws_per_batch=n_steps)
        out = {}
        for i in range(rank, n_batches, world):
            gen = st.advance_to(i)
            xT = torch.randn(shape, generator=gen)
            # (device == cpu here: the "device" default generator is the same global one manual_seed() returned, which is
            #  exactly the un-sharded reference's situation on a CPU run)
            zs = [torch.randn(shape) for _ in range(n_steps)]
            out[i] = (xT, zs)
        return out

    whole = run(0, 1)
    parts = {**run(0, 2), **run(1, 2)}
    assert sorted(parts) == sorted(whole) == list(range(n_batches))
    for i in whole:
        assert torch.equal(parts[i][0], whole[i][0])
        assert all(torch.equal(a, b) for a, b in zip(parts[i][1], whole[i][1]))
    assert not torch.equal(whole[0][0], whole[1][0])
    sch = types.SimpleNamespace(timesteps_host=torch.arange(11), kind="ddim_simple_orig", eta=0.85)
    exp = types.SimpleNamespace(scheduler=sch)
    assert IS.device_draws_per_batch(exp) == 10
    sch.eta = 0.0
    assert IS.device_draws_per_batch(exp) == 0 and IS.device_draws_per_batch(exp, new_eta=0.5) == 1
    sch.kind = "ddpm"
    assert IS.device_draws_per_batch(exp, sampling="project", max_T=4) == 4
