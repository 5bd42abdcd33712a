#!/bin/bash
O=gpurun_out; mkdir -p $O
for cfg in "8 50" "4 50" "4 75" "2 75" "8 75" "4 100"; do
set -- $cfg
NLC_SPLITK_MINCH=$1 NLC_SPLITK_FILL=$2 timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 2 --warmup 3 --batch 32 > $O/r02zt_b32.json 2>/dev/null
python - <<PY
import json
d=json.load(open('$O/r02zt_b32.json'))
print('c2 b32 minch=$1 fill=$2', round(d['value'],2), 'ms/ts', round(d['ms_per_timestep'],3))
PY
done
for cfg in "4 75" "4 100"; do
set -- $cfg
NLC_SPLITK_MINCH=$1 NLC_SPLITK_FILL=$2 timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 2 --warmup 3 > $O/r02zt_b256.json 2>/dev/null
python - <<PY
import json
d=json.load(open('$O/r02zt_b256.json'))
print('c2 b256 minch=$1 fill=$2', round(d['value'],2), 'ms/ts', round(d['ms_per_timestep'],3))
PY
done
