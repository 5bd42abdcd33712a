#!/bin/bash
O=gpurun_out; mkdir -p $O
show() { python -c "
import json,sys
try:
    d=json.load(open('$1'));print('$2',round(d['value'],2),'img/s',round(d['ms_per_timestep'],2),'ms/step conv',round(d['roofline']['frac'],3),'whole',round(d['roofline']['whole_step_frac'],3),'share',d['roofline']['conv_share_of_step'],d['clocks'])
except Exception as e: print('$2 FAILED',e)
"; }
python -m pytest tests/test_gpu_ops.py tests/test_gpu_nets.py tests/test_gpu_bench_arch.py tests/test_gpu_sampler.py -x -q -s > $O/r02e_pytest.log 2>&1; echo "pytest rc=$?"; grep -n "PSNR\|passed\|failed\|Error" $O/r02e_pytest.log | tail -30
for rep in 1 2; do
for cfgs in "fp16 0" "fp16 1" "bf16 0" "bf16 1"; do
set -- $cfgs
NLC_H16=$2 python bench.py --no-cpu-baseline --no-extras --precision $1 > $O/r02e_c2_$1_h$2_$rep.json 2> $O/r02e_c2_$1_h$2_$rep.err; show $O/r02e_c2_$1_h$2_$rep.json "c2 $1 h16=$2 rep$rep"
done; done
for cfgs in "fp16 0" "fp16 1" "bf16 0" "bf16 1"; do
set -- $cfgs
NLC_H16=$2 python bench.py --workload c5 --no-cpu-baseline --no-extras --precision $1 --steps 2 --warmup 2 > $O/r02e_c5_$1_h$2.json 2> $O/r02e_c5_$1_h$2.err; show $O/r02e_c5_$1_h$2.json "c5 $1 h16=$2"
done
NLC_H16=1 python bench.py --workload c3 --no-cpu-baseline --no-extras --precision fp16 --steps 2 --warmup 2 > $O/r02e_c3_fp16_h1.json 2> $O/r02e_c3.err; show $O/r02e_c3_fp16_h1.json "c3 fp16 h16=1"
