#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_bench_arch.py -q -k "dhariwal" > $O/r02j_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02j_pytest.log
timeout 300 python scripts/epi_ablate.py 1 > $O/r02j_epi_ablate.log 2>&1; echo "epi rc=$?"; cat $O/r02j_epi_ablate.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gn_apply_kernel --launch-skip 20 --launch-count 2 -f -o $O/r02j_ncu_gn python scripts/gn_bench.py 256 64 64 128 > $O/r02j_ncu_gn.log 2>&1; echo "ncu gn rc=$?"
ncu -i $O/r02j_ncu_gn.ncu-rep --page details --csv > $O/r02j_ncu_gn_details.csv 2>/dev/null
