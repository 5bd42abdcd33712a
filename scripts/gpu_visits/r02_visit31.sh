#!/bin/bash
# sustained (power-capped) A/B of this session's kernel changes inside bench.py's loop: old = staged epilogue + two-pass attention
O=gpurun_out; mkdir -p $O
for rep in 1 2; do
for cfg in "0 0" "1 1"; do
set -- $cfg
NLC_TMA_EPI=$1 NLC_ATTN_ONEPASS=$2 timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 2 --warmup 3 > $O/r02ze_bench_c2_t$1a$2_r$rep.json 2> $O/r02ze_bench_c2_t$1a$2_r$rep.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open('$O/r02ze_bench_c2_t$1a$2_r$rep.json'))
print('c2 tma$1 attn$2 rep$rep', round(d['value'],2), 'ms/ts', round(d['ms_per_timestep'],3), 'conv', round(d['roofline']['frac'],3), 'whole', round(d['roofline']['whole_step_frac'],3), d['clocks'])
PY
done
done
for cfg in "0 0" "1 1"; do
set -- $cfg
NLC_TMA_EPI=$1 NLC_ATTN_ONEPASS=$2 timeout 600 python bench.py --workload c5 --no-extras --no-cpu-baseline --steps 2 --warmup 3 > $O/r02ze_bench_c5_t$1a$2.json 2> $O/r02ze_bench_c5_t$1a$2.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open('$O/r02ze_bench_c5_t$1a$2.json'))
print('c5 tma$1 attn$2', round(d['value'],2), 'ms/ts', round(d['ms_per_timestep'],3), 'conv', round(d['roofline']['frac'],3), 'whole', round(d['roofline']['whole_step_frac'],3), d['clocks'])
PY
done
