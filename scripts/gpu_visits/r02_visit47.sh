#!/bin/bash
# full validation of the session's state: every GPU test, smoke(), the driver's default bench line, the reference arm
O=gpurun_out; mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q -x > $O/r02zp_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02zp_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02zp_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r02zp_smoke.log
SECONDS=0
timeout 1500 python bench.py > $O/r02zp_bench.json 2> $O/r02zp_bench.err; echo "bench rc=$? wall ${SECONDS}s"
python - <<PY
import json
d=json.load(open('$O/r02zp_bench.json'))
print('c2', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'ms/ts', round(d['ms_per_timestep'],3), 'conv', round(d['roofline']['frac'],3), 'whole', round(d['roofline']['whole_step_frac'],3), d['clocks'], d['roofline'].get('traffic'))
a=d['adm256']; print('adm256', round(a['value'],2), 'e2e', round(a['e2e']['value'],2), a['ms_per_timestep'], a['roofline']['frac'], a['roofline']['whole_step_frac'])
print('tf32', d['tf32']['value'], 'bf16', d['bf16']['value'], 'k2b', d['kernel_to_beat'], 'cpu', d['cpu_baseline'])
PY
tail -3 $O/r02zp_bench.err
SECONDS=0
timeout 900 python bench.py --impl reference --steps 1 --warmup 1 > $O/r02zp_bench_ref.json 2> $O/r02zp_bench_ref.err; echo "ref rc=$? wall ${SECONDS}s"; cat $O/r02zp_bench_ref.json | cut -c1-400
