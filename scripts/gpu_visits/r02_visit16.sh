#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_training.py tests/test_gpu_nets.py tests/test_gpu_image_sample.py -q -x > $O/r02p_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 $O/r02p_pytest.log
timeout 300 python bench.py --batch 128 --no-extras --no-cpu-baseline --steps 2 --warmup 2 > $O/r02p_b128_graph.json 2>$O/r02p_b128_graph.err; python -c "
import json;d=json.load(open('$O/r02p_b128_graph.json'));print('B128 graph', d['value'], d['ms_per_timestep'])"
timeout 300 python bench.py --batch 128 --no-extras --no-cpu-baseline --steps 2 --warmup 2 --no-graph > $O/r02p_b128_nograph.json 2>$O/r02p_b128_nograph.err; python -c "
import json;d=json.load(open('$O/r02p_b128_nograph.json'));print('B128 eager', d['value'], d['ms_per_timestep'])"
