#!/bin/bash
O=gpurun_out; mkdir -p $O
show() { python -c "
import json,sys
try:
    d=json.load(open('$1'));print('$2',round(d['value'],2),'img/s',round(d['ms_per_timestep'],2),'ms/step conv',round(d['roofline']['frac'],3),'whole',round(d['roofline']['whole_step_frac'],3),'share',round(d['roofline']['conv_share_of_step'] or 0,3),d['clocks'])
except Exception as e: print('$2 FAILED',e)
"; }
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "slab or 16bit" > $O/r02f_pytest_slab.log 2>&1; echo "pytest slab rc=$?"; tail -15 $O/r02f_pytest_slab.log
for m in 0 2; do timeout 300 python scripts/epi_ablate.py $m > $O/r02f_epi_slab$m.log 2>&1; echo "epi slab=$m rc=$?"; cat $O/r02f_epi_slab$m.log; done
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_nets.py tests/test_gpu_bench_arch.py tests/test_gpu_sampler.py -x -q -s > $O/r02f_pytest.log 2>&1; echo "pytest rc=$?"; grep -n "PSNR\|passed\|failed\|Error" $O/r02f_pytest.log | tail -30
for cfgs in "fp16 0 0" "fp16 1 0" "fp16 1 1" "fp16 1 2" "bf16 0 0" "bf16 1 2"; do
set -- $cfgs
NLC_H16=$2 NLC_SLAB=$3 timeout 600 python bench.py --no-cpu-baseline --no-extras --precision $1 > $O/r02f_c2_$1_h$2_s$3.json 2> $O/r02f_c2_$1_h$2_s$3.err; show $O/r02f_c2_$1_h$2_s$3.json "c2 $1 h16=$2 slab=$3"
done
for cfgs in "fp16 0 0" "fp16 1 0" "fp16 1 2" "bf16 0 0" "bf16 1 0"; do
set -- $cfgs
NLC_H16=$2 NLC_SLAB=$3 timeout 600 python bench.py --workload c5 --no-cpu-baseline --no-extras --precision $1 --steps 2 --warmup 2 > $O/r02f_c5_$1_h$2_s$3.json 2> $O/r02f_c5_$1_h$2_s$3.err; show $O/r02f_c5_$1_h$2_s$3.json "c5 $1 h16=$2 slab=$3"
done
for s in 0 2; do
NLC_H16=1 NLC_SLAB=$s timeout 600 python bench.py --workload c3 --no-cpu-baseline --no-extras --precision fp16 --steps 2 --warmup 2 > $O/r02f_c3_fp16_h1_s$s.json 2> $O/r02f_c3_s$s.err; show $O/r02f_c3_fp16_h1_s$s.json "c3 fp16 h16=1 slab=$s"
done
