#!/bin/bash
# GroupNorm apply: one MUFU.RCP per four SiLU elements (NLC_GN_RCP4=0 is the control): parity, isolated bandwidth, burst step, sustained loop
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ops.py -q -x -k "groupnorm" > $O/r02zf_pytest_gn.log 2>&1; echo "pytest gn rc=$?"; tail -3 $O/r02zf_pytest_gn.log
for m in 1 0; do
NLC_GN_RCP4=$m timeout 300 python scripts/gn_bench.py > $O/r02zf_gn_bench_r$m.log 2>&1; echo "gn_bench rcp4=$m"; cat $O/r02zf_gn_bench_r$m.log
done
for m in 1 0; do
NLC_GN_RCP4=$m timeout 600 python scripts/step_profile.py c2 256 fp16 > $O/r02zf_step_c2_fp16_r$m.log 2>&1; head -3 $O/r02zf_step_c2_fp16_r$m.log | tail -2; grep "groupnorm 64x64 C128\|groupnorm 32x32 C256" $O/r02zf_step_c2_fp16_r$m.log
done
for rep in 1 2; do
for m in 0 1; do
NLC_GN_RCP4=$m timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 2 --warmup 3 > $O/r02zf_bench_c2_r${m}_rep$rep.json 2> $O/r02zf_bench_c2_r${m}_rep$rep.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open('$O/r02zf_bench_c2_r${m}_rep$rep.json'))
print('c2 rcp4=$m rep$rep', round(d['value'],2), 'ms/ts', round(d['ms_per_timestep'],3), 'conv', round(d['roofline']['frac'],3), 'whole', round(d['roofline']['whole_step_frac'],3), d['clocks'])
PY
done
done
for m in 0 1; do
NLC_GN_RCP4=$m timeout 600 python bench.py --workload c5 --no-extras --no-cpu-baseline --steps 2 --warmup 3 > $O/r02zf_bench_c5_r$m.json 2> $O/r02zf_bench_c5_r$m.err; echo "rc=$?"
python - <<PY
import json
d=json.load(open('$O/r02zf_bench_c5_r$m.json'))
print('c5 rcp4=$m', round(d['value'],2), 'ms/ts', round(d['ms_per_timestep'],3), 'conv', round(d['roofline']['frac'],3), 'whole', round(d['roofline']['whole_step_frac'],3), d['clocks'])
PY
done
