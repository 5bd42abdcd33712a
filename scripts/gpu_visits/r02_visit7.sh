#!/bin/bash
O=gpurun_out; mkdir -p $O
for cfgs in "fp16 0" "bf16 0" "fp16 1"; do
set -- $cfgs
NLC_H16=$2 timeout 600 python scripts/step_profile.py c2 256 $1 > $O/r02g_step_c2_$1_h$2.log 2>&1; echo "c2 $1 h16=$2"; head -28 $O/r02g_step_c2_$1_h$2.log
done
for cfgs in "fp16 0" "bf16 0"; do
set -- $cfgs
NLC_H16=$2 timeout 600 python scripts/step_profile.py adm256 16 $1 > $O/r02g_step_adm_$1_h$2.log 2>&1; echo "adm $1 h16=$2"; head -24 $O/r02g_step_adm_$1_h$2.log
done
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_nets.py tests/test_gpu_bench_arch.py tests/test_gpu_sampler.py tests/test_gpu_edm.py -x -q -s > $O/r02g_pytest.log 2>&1; echo "pytest rc=$?"; grep -n "PSNR\|passed\|failed\|Error" $O/r02g_pytest.log | tail -30
