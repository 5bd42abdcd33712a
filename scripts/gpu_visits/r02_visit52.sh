#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q -x > $O/r02zu_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02zu_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02zu_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r02zu_smoke.log
SECONDS=0
timeout 1500 python bench.py > $O/r02zu_bench.json 2> $O/r02zu_bench.err; echo "bench rc=$? wall ${SECONDS}s"
python - <<PY
import json
d=json.load(open('$O/r02zu_bench.json'))
print('c2', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'ms/ts', round(d['ms_per_timestep'],3), 'conv', round(d['roofline']['frac'],3), 'whole', round(d['roofline']['whole_step_frac'],3), d['clocks'])
a=d['adm256']; print('adm256', round(a['value'],2), 'e2e', round(a['e2e']['value'],2), a['ms_per_timestep'], a['roofline']['frac'], a['roofline']['whole_step_frac'])
print('tf32', d['tf32']['value'], 'bf16', d['bf16']['value'])
PY
tail -2 $O/r02zu_bench.err
timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 2 --warmup 3 --batch 32 > $O/r02zu_bench_b32.json 2>/dev/null
python - <<PY
import json
d=json.load(open('$O/r02zu_bench_b32.json'))
print('c2 b32', round(d['value'],2), 'ms/ts', round(d['ms_per_timestep'],3))
PY
