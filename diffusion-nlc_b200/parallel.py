"""Batch sharding of the sampling loop across GPUs: samples are independent, so every rank runs the
whole loop on its slice of the global batch; the only collectives are the all-gather of finished images / metric
partial sums at the end of a batch and, optionally, a 3-scalar all-reduce per step that reproduces the reference's
batch-global decisions (mean constraint loss for best-x0 selection, NaN flag, max t; src/experiments.py:371-389,
image_sample.py:471).  One process per GPU (torchrun), NCCL on GPUs, gloo for the CPU-side tests."""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(global_batch, rank, world_size):
    """Rows [lo, hi) of the global batch owned by `rank` (contiguous, remainder spread over the first ranks)."""
    base, rem = divmod(global_batch, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def sharded_noise(global_shape, seed, rank, world_size, n_draws=1):
    """The reference draws x_T (and, on the CPU, every step's noise) for the *whole* batch from one generator
    (src/experiments.py:260-271).  Each rank replays that stream and keeps its rows, so a sharded run reproduces
    the un-sharded one sample for sample.  Returns `n_draws` tensors of shape [hi-lo, ...]."""
    lo, hi = shard_range(global_shape[0], rank, world_size)
    gen = torch.Generator().manual_seed(seed)
    return [torch.randn(global_shape, generator=gen)[lo:hi].clone() for _ in range(n_draws)]


def gather_images(local, global_batch=None):
    """All-gather finished images along the batch dimension (ragged shards allowed)."""
    rank, ws = world()
    if ws == 1:
        return local
    sizes = [shard_range(global_batch, r, ws) for r in range(ws)] if global_batch is not None else None
    if sizes is None or len({hi - lo for lo, hi in sizes}) == 1:
        out = [torch.empty_like(local) for _ in range(ws)]
        dist.all_gather(out, local.contiguous())
        return torch.cat(out, dim=0)
    n_max = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((n_max,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(ws)]
    dist.all_gather(out, pad)
    return torch.cat([o[:hi - lo] for o, (lo, hi) in zip(out, sizes)], dim=0)


def global_step_scalars(loss_sum, nan_flag, t_max):
    """(sum of per-sample constraint losses, any NaN, max t) over all ranks: one 3-float all-reduce pair."""
    rank, ws = world()
    if ws == 1:
        return loss_sum, nan_flag, t_max
    s = torch.stack([torch.as_tensor(loss_sum, dtype=torch.float32), torch.as_tensor(nan_flag, dtype=torch.float32)])
    dist.all_reduce(s, op=dist.ReduceOp.SUM)
    m = torch.as_tensor(t_max, dtype=torch.float32).reshape(1).clone()
    dist.all_reduce(m, op=dist.ReduceOp.MAX)
    return s[0], s[1] > 0, m[0]


def reduce_metric_sums(sums):
    """Sum metric partials (e.g. squared error, L1 residuals, counts) over ranks."""
    rank, ws = world()
    if ws == 1:
        return sums
    t = sums.clone()
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t
