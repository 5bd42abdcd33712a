#!/bin/bash
O=gpurun_out; mkdir -p $O
python -m pytest tests/test_gpu_bench_arch.py -x -q -s > $O/r02c_pytest_bench_arch.log 2>&1; echo "pytest bench_arch rc=$?"; tail -25 $O/r02c_pytest_bench_arch.log
python scripts/epi_ablate.py > $O/r02c_epi_ablate.log 2>&1; echo "epi rc=$?"; cat $O/r02c_epi_ablate.log
python scripts/prof_conv.py > $O/r02c_prof_conv.log 2>&1; cat $O/r02c_prof_conv.log
ncu --set full --clock-control none --import-source on -k regex:conv_tc -s 6 -c 1 -f -o $O/r02c_ncu_conv_c2 python scripts/prof_conv.py > $O/r02c_ncu_conv.log 2>&1; echo "ncu rc=$?"; tail -3 $O/r02c_ncu_conv.log
