#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 python scripts/gn_bench.py > $O/r02k_gn_bench.log 2>&1; echo "gn rc=$?"; cat $O/r02k_gn_bench.log
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_nets.py tests/test_gpu_adm.py tests/test_gpu_bench_arch.py -q -s > $O/r02k_pytest.log 2>&1; echo "pytest rc=$?"; grep -n "PSNR\|control\|passed\|failed\|Error\|FAILED" $O/r02k_pytest.log | tail -20
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_ -f -o $O/r02k_ncu_epi python scripts/epi_ncu.py > $O/r02k_ncu_epi.log 2>&1; echo "ncu epi rc=$?"
ncu -i $O/r02k_ncu_epi.ncu-rep --page details --csv > $O/r02k_ncu_epi_details.csv 2>/dev/null
ncu -i $O/r02k_ncu_epi.ncu-rep --page raw --csv > $O/r02k_ncu_epi_raw.csv 2>/dev/null
timeout 600 python scripts/step_profile.py c2 256 fp16 > $O/r02k_step_c2_fp16.log 2>&1; head -12 $O/r02k_step_c2_fp16.log
