#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_bench_arch.py tests/test_gpu_nets.py tests/test_gpu_image_sample.py tests/test_gpu_constrained.py tests/test_gpu_sampler.py -q > $O/r02q_pytest.log 2>&1; echo "pytest rc=$?"; grep -n "passed\|failed\|Error\|FAILED" $O/r02q_pytest.log | tail -20
for b in 128 256; do
timeout 300 python bench.py --batch $b --no-extras --no-cpu-baseline --steps 3 --warmup 3 > $O/r02q_b${b}_graph.json 2>$O/r02q_b${b}_graph.err; python -c "
import json;d=json.load(open('$O/r02q_b${b}_graph.json'));print('B$b graph', d['value'], d['ms_per_timestep'], 'e2e', d['e2e']['value'])"
done
timeout 600 python bench.py --workload c5 --no-extras --no-cpu-baseline --steps 2 --warmup 2 > $O/r02q_c5.json 2>$O/r02q_c5.err; python -c "
import json;d=json.load(open('$O/r02q_c5.json'));print('c5', d['value'], d['ms_per_timestep'], 'e2e', d['e2e']['value'])"
