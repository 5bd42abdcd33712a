#!/bin/bash
# One GPU-box visit: parity tests, smoke, bench (both arms), ncu launch list of the bench command.
# Usage (from the container): gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh TAG'
TAG=${1:-r01}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/${TAG}_pytest_gpu.log
python __graft_entry__.py smoke > gpurun_out/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/${TAG}_smoke.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cat gpurun_out/${TAG}_bench.json
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; echo "ref rc=$?"
cat gpurun_out/${TAG}_bench_reference.json
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2100 -c 2200 --csv \
    --log-file gpurun_out/${TAG}_launches_c2.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline \
    > gpurun_out/${TAG}_ncu_run.log 2>&1; echo "ncu rc=$?"
