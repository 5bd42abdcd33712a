"""Sharded == un-sharded for the batch-global decisions of the constrained loop (src/experiments.py:371-389:
the best-x0 step is chosen by the BATCH-MEAN constraint loss, image_sample.py:471 likewise).  Run under torchrun on N GPUs:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        scripts/exact_global_check.py

Every rank runs (a) the whole batch alone (the reference's behaviour), (b) its shard with `exact_global=True` - the per-step
loss sum all-reduced over NCCL inside the captured CUDA graph - and (c) its shard deciding on its local mean.  (b) must
reproduce (a) row for row; (c) is reported (it may legitimately pick another step)."""
import os
import sys

local = int(os.environ.get("LOCAL_RANK", "0"))
os.environ["CUDA_VISIBLE_DEVICES"] = os.environ.get("CUDA_VISIBLE_DEVICES", "0,1,2,3,4,5,6,7").split(",")[local]

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from functools import partial  # noqa: E402

from test_gpu_constrained import TASKS, _setup  # noqa: E402


def main():
    dist.init_process_group("nccl", device_id=torch.device("cuda:0"))
    rank, world = dist.get_rank(), dist.get_world_size()
    dev = torch.device("cuda:0")
    golden = torch.load(os.path.join(ROOT, "tests", "golden", "loops3_constrained.pt"), weights_only=True)
    ok = True
    for prec in ("fp32", "fp16"):
        for key in TASKS:
            exp, sch, con, case, y, cfn, closs = _setup(prec, golden, key)
            B = case["x_true"].shape[0]
            assert B % world == 0
            per = B // world
            lo, hi = rank * per, (rank + 1) * per
            shape = tuple(case["z"].shape)
            xT = (case["z"] / (1 / (case["sigmas"][0] ** 2 + 1)).sqrt()).to(dev)
            kw = dict(style="pred", norm_eps=True, refine_prior_sigma=True, return_log=False, chunk_size=1,
                      sigma_pred_threshold=960, to_cpu=False, graph=True)
            full, _ = exp.denoise_loop(shape=shape, xT=xT, constrain_fn=cfn, constrain_loss=closs,
                                       noise_fn=lambda i, like: case["noises"][i].to(dev), **kw)
            ys = y[lo:hi]
            cfn_s = partial(con.constraint_fn, y=ys, lambda_t=con.lr)
            closs_s = partial(con.loss, y=ys)
            res = {}
            for exact in (True, False):
                out, _ = exp.denoise_loop(shape=(per,) + shape[1:], xT=xT[lo:hi], constrain_fn=cfn_s, constrain_loss=closs_s,
                                          noise_fn=lambda i, like: case["noises"][i][lo:hi].to(dev), exact_global=exact, **kw)
                parts = [torch.empty_like(out) for _ in range(world)]
                dist.all_gather(parts, out.contiguous())
                res[exact] = torch.cat(parts)
            d_exact = (res[True] - full).abs().max().item()
            d_local = (res[False] - full).abs().max().item()
            ok = ok and d_exact < 1e-6
            if rank == 0:
                print("%-5s %-22s %d ranks: exact_global max|sharded - unsharded| = %.2e; local-mean decision: %.2e" % (
                    prec, key, world, d_exact, d_local), flush=True)
    flag = torch.tensor([1.0 if ok else 0.0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("RESULT", "ok" if flag.item() == 1.0 else "MISMATCH", flush=True)
    dist.destroy_process_group()
    sys.exit(0 if flag.item() == 1.0 else 1)


if __name__ == "__main__":
    main()
