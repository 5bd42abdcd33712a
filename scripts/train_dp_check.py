"""Data-parallel check of the training slice: under torchrun with N ranks, every rank runs
`SigmaTrainer.step()` on its own shard; afterwards all ranks must hold identical parameters, equal to what one process
gets from the MEAN of the ranks' gradients with torch.optim.AdamW (the reference's DDP runs under no_sync() and would
leave the ranks diverged).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 scripts/train_dp_check.py
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import nlc_b200  # noqa: E402,F401
from nlc_b200 import training as T  # noqa: E402


def head():
    torch.manual_seed(0)
    return torch.nn.Sequential(torch.nn.Conv2d(16, 8, 3, padding=1), torch.nn.SiLU(), torch.nn.Flatten(),
                               torch.nn.Linear(8 * 4 * 4, 1))


def shard(rank):
    g = torch.Generator().manual_seed(100 + rank)
    return torch.randn(6, 16, 4, 4, generator=g), torch.rand(6, 1, generator=g) + 0.5


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    net = head().to(dev)
    tr = T.SigmaTrainer(net, lr=1e-2, weight_decay=0.01, ema_rate=0.9)
    loss_fn = torch.nn.MSELoss()
    for it in range(3):
        x, y = shard(rank * 10 + it)
        tr.zero_grad()
        loss_fn(net(x.to(dev)) + 1, y.to(dev)).backward()
        tr.step()
    flat = tr.flat.clone()
    gathered = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(gathered, flat)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    ok = same
    if rank == 0:
        ref = head()
        opt = torch.optim.AdamW(ref.parameters(), lr=1e-2, weight_decay=0.01)
        for it in range(3):
            opt.zero_grad()
            for r in range(world):  # the mean over ranks of the per-rank mean losses
                x, y = shard(r * 10 + it)
                (loss_fn(ref(x) + 1, y) / world).backward()
            opt.step()
        err = max((a.detach().cpu() - b.detach()).abs().max().item() for a, b in zip(net.parameters(), ref.parameters()))
        ok = ok and err < 1e-5
        print("train_dp_check: world %d, ranks identical %s, max |param - single-process AdamW on the mean gradient| %.2e -> %s"
              % (world, same, err, "OK" if ok else "FAIL"), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
