#!/bin/bash
# final-state evidence: ncu launch list of the bench command, --set full captures of the dominant kernels inside it
O=gpurun_out; mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2100 -c 2200 --csv --log-file $O/r02z_launches_c2_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $O/r02z_ncu_run.log 2>&1; echo "ncu launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_slab_kernel --launch-skip 300 --launch-count 6 -f -o $O/r02z_ncu_conv_c2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $O/r02z_ncu_conv_c2.log 2>&1; echo "ncu c2 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip 1 --launch-count 6 -f -o $O/r02z_ncu_conv_c5 python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > $O/r02z_ncu_conv_c5.log 2>&1; echo "ncu c5 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_fused1 --launch-skip 20 --launch-count 3 -f -o $O/r02z_ncu_attn_c5 python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > $O/r02z_ncu_attn_c5.log 2>&1; echo "ncu attn rc=$?"
ls -la $O/r02z_*
