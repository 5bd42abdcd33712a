#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_ops.py tests/test_gpu_nets.py tests/test_gpu_adm.py tests/test_gpu_edm.py tests/test_gpu_bench_arch.py tests/test_gpu_fp32_mode.py tests/test_gpu_canaries.py -q -x > $O/r02l_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r02l_pytest.log
timeout 300 python scripts/epi_ablate.py 1 > $O/r02l_epi_ablate.log 2>&1; echo "epi rc=$?"; cat $O/r02l_epi_ablate.log
timeout 600 python scripts/step_profile.py c2 256 fp16 > $O/r02l_step_c2_fp16.log 2>&1; head -30 $O/r02l_step_c2_fp16.log
timeout 600 python scripts/step_profile.py adm256 16 fp16 > $O/r02l_step_adm_fp16.log 2>&1; head -24 $O/r02l_step_adm_fp16.log
