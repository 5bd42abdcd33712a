#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/exact_global_check.py > $O/r02o_exact_global_n2.log 2>&1; echo "exact_global rc=$?"; grep -v "^W\|^\[W\|NCCL" $O/r02o_exact_global_n2.log | tail -12
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 --no-cpu-baseline > $O/r02o_bench_n2.json 2> $O/r02o_bench_n2.err; echo "bench n2 rc=$?"; tail -c 1500 $O/r02o_bench_n2.json; tail -3 $O/r02o_bench_n2.err
