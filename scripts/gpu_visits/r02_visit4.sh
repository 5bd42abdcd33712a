#!/bin/bash
O=gpurun_out; mkdir -p $O
python scripts/epi_ablate.py > $O/r02d_epi_ablate.log 2>&1; echo "epi rc=$?"; cat $O/r02d_epi_ablate.log
python -m pytest tests/test_gpu_ops.py tests/test_gpu_bench_arch.py tests/test_gpu_sampler.py -x -q -s > $O/r02d_pytest.log 2>&1; echo "pytest rc=$?"; grep -n "PSNR\|passed\|failed\|Error" $O/r02d_pytest.log | tail -30
for h in 1 0; do
NLC_H16=$h python bench.py --no-cpu-baseline --no-extras --precision fp16 --steps 2 --warmup 2 > $O/r02d_bench_c2_fp16_h$h.json 2> $O/r02d_bench_c2_fp16_h$h.err; echo "bench h16=$h rc=$?"; python -c "
import json;d=json.load(open('$O/r02d_bench_c2_fp16_h$h.json'));print(d['value'],d['ms_per_timestep'],d['roofline']['frac'],d['roofline']['whole_step_frac'],d['clocks'])"
done
NLC_H16=1 python bench.py --no-cpu-baseline --no-extras --precision bf16 --steps 2 --warmup 2 > $O/r02d_bench_c2_bf16_h1.json 2>/dev/null; python -c "
import json;d=json.load(open('$O/r02d_bench_c2_bf16_h1.json'));print('bf16 h16',d['value'],d['ms_per_timestep'],d['roofline']['frac'],d['clocks'])"
for h in 1 0; do
NLC_H16=$h python bench.py --workload c5 --no-cpu-baseline --no-extras --precision fp16 --steps 2 --warmup 2 > $O/r02d_bench_c5_fp16_h$h.json 2> $O/r02d_bench_c5_fp16_h$h.err; echo "bench c5 h16=$h rc=$?"; python -c "
import json;d=json.load(open('$O/r02d_bench_c5_fp16_h$h.json'));print(d['value'],d['ms_per_timestep'],d['roofline']['frac'],d['roofline']['whole_step_frac'],d['clocks'])"
done
python bench.py --workload c5 --no-cpu-baseline --no-extras --precision bf16 --steps 2 --warmup 2 > $O/r02d_bench_c5_bf16.json 2>/dev/null; python -c "
import json;d=json.load(open('$O/r02d_bench_c5_bf16.json'));print('c5 bf16',d['value'],d['ms_per_timestep'],d['roofline']['frac'],d['roofline']['whole_step_frac'],d['clocks'])"
