#!/bin/bash
# does the output tensor map's L2 promotion cause the extra DRAM reads of the TMA-store epilogue?  ncu DRAM bytes + sustained c5 / c2
O=gpurun_out; mkdir -p $O
for pr in 0 1; do
NLC_TMA_OUT_PROMO=$pr timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv_tc_kernel --launch-skip 1 --launch-count 4 --csv --log-file $O/r02zk_dram_c5_promo$pr.csv python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > /dev/null 2>&1; echo "ncu promo$pr rc=$?"
grep -E "dram__bytes|gpu__time" $O/r02zk_dram_c5_promo$pr.csv | awk -F'","' '{print $5, $(NF-2), $(NF-1), $NF}' | head -12
done
for rep in 1 2; do
for pr in 1 0; do
NLC_TMA_OUT_PROMO=$pr timeout 600 python bench.py --workload c5 --no-extras --no-cpu-baseline --steps 2 --warmup 3 > $O/r02zk_bench_c5_promo${pr}_$rep.json 2>/dev/null
python - <<PY
import json
d=json.load(open('$O/r02zk_bench_c5_promo${pr}_$rep.json'))
print('c5 promo$pr rep$rep', round(d['value'],2), 'ms/ts', round(d['ms_per_timestep'],3), 'conv', round(d['roofline']['frac'],3), 'whole', round(d['roofline']['whole_step_frac'],3), d['clocks'])
PY
done
done
for pr in 1 0; do
NLC_TMA_OUT_PROMO=$pr timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 2 --warmup 3 > $O/r02zk_bench_c2_promo$pr.json 2>/dev/null
python - <<PY
import json
d=json.load(open('$O/r02zk_bench_c2_promo$pr.json'))
print('c2 promo$pr', round(d['value'],2), 'ms/ts', round(d['ms_per_timestep'],3), 'conv', round(d['roofline']['frac'],3), 'whole', round(d['roofline']['whole_step_frac'],3), d['clocks'])
PY
done
