#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_nets.py tests/test_gpu_adm.py tests/test_gpu_edm.py tests/test_gpu_bench_arch.py tests/test_gpu_training.py tests/test_gpu_canaries.py -q > $O/r02s_pytest.log 2>&1; echo "pytest rc=$?"; grep -n "passed\|failed\|FAILED" $O/r02s_pytest.log | tail -8
timeout 300 python scripts/epi_ablate.py 1 > $O/r02s_epi_ablate.log 2>&1; echo "epi rc=$?"; head -17 $O/r02s_epi_ablate.log
timeout 600 python scripts/step_profile.py c2 256 fp16 > $O/r02s_step_c2_fp16.log 2>&1; head -8 $O/r02s_step_c2_fp16.log
timeout 600 python scripts/step_profile.py adm256 16 fp16 > $O/r02s_step_adm_fp16.log 2>&1; head -8 $O/r02s_step_adm_fp16.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip 1 --launch-count 4 -f -o $O/r02s_ncu_conv_c5 python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > $O/r02s_ncu_conv_c5.log 2>&1; echo "ncu c5 rc=$?"
