#!/bin/bash
# programmatic dependent launch (griddepcontrol) A/B inside the CUDA-graph replay, + parity of the affected kernels
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_nets.py tests/test_gpu_sampler.py tests/test_gpu_adm.py -q -x > $O/r02zd_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02zd_pytest.log
for pdl in 1 0; do
for b in 256 32; do
NLC_PDL=$pdl timeout 600 python bench.py --no-extras --no-cpu-baseline --steps 2 --warmup 3 --batch $b > $O/r02zd_bench_c2_b${b}_pdl$pdl.json 2> $O/r02zd_bench_c2_b${b}_pdl$pdl.err; echo "bench b$b pdl$pdl rc=$?"
python - <<PY
import json
d=json.load(open('$O/r02zd_bench_c2_b${b}_pdl$pdl.json'))
print('c2 b$b pdl$pdl', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'ms/ts', round(d['ms_per_timestep'],3), 'conv', round(d['roofline']['frac'],3), 'whole', round(d['roofline']['whole_step_frac'],3))
PY
done
NLC_PDL=$pdl timeout 600 python bench.py --workload c5 --no-extras --no-cpu-baseline --steps 2 --warmup 3 > $O/r02zd_bench_c5_pdl$pdl.json 2> $O/r02zd_bench_c5_pdl$pdl.err; echo "bench c5 pdl$pdl rc=$?"
python - <<PY
import json
d=json.load(open('$O/r02zd_bench_c5_pdl$pdl.json'))
print('c5 pdl$pdl', round(d['value'],2), 'e2e', round(d['e2e']['value'],2), 'ms/ts', round(d['ms_per_timestep'],3), 'conv', round(d['roofline']['frac'],3), 'whole', round(d['roofline']['whole_step_frac'],3))
PY
done
