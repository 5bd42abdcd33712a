#!/bin/bash
O=gpurun_out; mkdir -p $O
for w in c3 c4 c4p c5 c5cs; do
  timeout 900 python bench.py --workload $w --steps 2 --warmup 3 --no-extras > $O/r02zm_bench_$w.json 2> $O/r02zm_bench_$w.err; echo "bench $w rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('$O/r02zm_bench_$w.json'))
    print('$w', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'ms/ts', round(d['ms_per_timestep'],2), 'conv', round(d['roofline']['frac'],3), 'whole', round(d['roofline']['whole_step_frac'],3), d['dtype'])
except Exception as e: print('$w FAILED', e)
PY
  tail -2 $O/r02zm_bench_$w.err
done
