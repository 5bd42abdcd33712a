#!/bin/bash
# split-K A/B inside the CUDA-graph replay (bench.py), where the small launches are GPU-bound
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ops.py -q -x > $O/r02zs_pytest_ops.log 2>&1; echo "pytest ops rc=$?"; tail -3 $O/r02zs_pytest_ops.log
for b in 32 256; do
for m in 1 0 1 0; do
NLC_SPLITK=$m timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 2 --warmup 3 --batch $b > $O/r02zs_bench_c2_b${b}_k$m.json 2>/dev/null
python - <<PY
import json
d=json.load(open('$O/r02zs_bench_c2_b${b}_k$m.json'))
print('c2 b$b splitk=$m', round(d['value'],2), 'ms/ts', round(d['ms_per_timestep'],3), 'conv', round(d['roofline']['frac'],3), d['clocks']['sm_mhz'])
PY
done
done
