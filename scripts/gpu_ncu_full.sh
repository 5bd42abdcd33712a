#!/bin/bash
# ncu --set full captures of the dominant kernels inside the bench command (c2 workload).
# Usage: gpurun --timeout 1200 -- 'bash scripts/gpu_ncu_full.sh TAG'
TAG=${1:-r01}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip 700 --launch-count 6 \
    -f -o gpurun_out/${TAG}_ncu_conv_c2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_conv.log 2>&1
echo "ncu conv rc=$?"
ncu --set full --clock-control none --import-source on -k regex:gn_apply_kernel --launch-skip 400 --launch-count 6 \
    -f -o gpurun_out/${TAG}_ncu_gn_c2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_gn.log 2>&1
echo "ncu gn rc=$?"
