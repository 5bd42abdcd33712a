#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 python scripts/gn_bench.py > $O/r02i_gn_bench.log 2>&1; echo "gn rc=$?"; cat $O/r02i_gn_bench.log
timeout 1500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_nets.py tests/test_gpu_adm.py tests/test_gpu_edm.py tests/test_gpu_bench_arch.py tests/test_gpu_sampler.py tests/test_gpu_constrained.py -q -s > $O/r02i_pytest.log 2>&1; echo "pytest rc=$?"; grep -n "PSNR\|passed\|failed\|Error\|FAILED" $O/r02i_pytest.log | tail -40
for r in 0 1; do
NLC_R16=$r timeout 600 python scripts/step_profile.py c2 256 fp16 > $O/r02i_step_c2_fp16_r$r.log 2>&1; echo "c2 fp16 r16=$r"; head -24 $O/r02i_step_c2_fp16_r$r.log
done
for r in 0 1; do
NLC_R16=$r timeout 600 python scripts/step_profile.py adm256 16 fp16 > $O/r02i_step_adm_fp16_r$r.log 2>&1; echo "adm fp16 r16=$r"; head -20 $O/r02i_step_adm_fp16_r$r.log
done
