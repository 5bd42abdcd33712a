#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29547 bench.py --gpus 8 --steps 2 --warmup 3 --no-cpu-baseline > $O/r02zv_bench_n8.json 2> $O/r02zv_bench_n8.err; echo "bench n8 rc=$?"; python - <<PY
import json
d=json.load(open('$O/r02zv_bench_n8.json'))
print("N8 c2 weak", d["value"], "e2e", d["e2e"]["value"], "ms/ts", d["ms_per_timestep"], d["clocks"])
print("strong", d.get("strong"))
a=d["adm256"]; print("adm256", a["value"], a["ms_per_timestep"], a["e2e"]["value"], a["roofline"]["frac"], a["roofline"]["whole_step_frac"], a["clocks"])
PY
tail -3 $O/r02zv_bench_n8.err
