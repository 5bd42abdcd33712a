#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_training.py -q > $O/r02r_pytest_training.log 2>&1; echo "pytest training rc=$?"; tail -3 $O/r02r_pytest_training.log
# compute-sanitizer memcheck on the smallest end-to-end paths
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02r_sanitizer_smoke.log 2>&1; echo "sanitizer smoke rc=$?"; tail -4 $O/r02r_sanitizer_smoke.log
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_fid.py tests/test_gpu_training.py -q -k "preprocess or im2col or pooling or statistics or native_sigma_model_training" > $O/r02r_sanitizer_fid_train.log 2>&1; echo "sanitizer fid/train rc=$?"; tail -4 $O/r02r_sanitizer_fid_train.log
# launch list of the bench command
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 2100 -c 2200 --csv --log-file $O/r02r_launches_c2_bench.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $O/r02r_ncu_run.log 2>&1; echo "ncu launch list rc=$?"
# full captures of the dominant conv kernels inside the bench commands
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_slab_kernel --launch-skip 300 --launch-count 6 -f -o $O/r02r_ncu_conv_c2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $O/r02r_ncu_conv_c2.log 2>&1; echo "ncu c2 rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_tc_kernel --launch-skip 260 --launch-count 10 -f -o $O/r02r_ncu_conv_c5 python bench.py --workload c5 --steps 1 --warmup 1 --no-cpu-baseline --no-extras > $O/r02r_ncu_conv_c5.log 2>&1; echo "ncu c5 rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gn_apply_lean --launch-skip 200 --launch-count 4 -f -o $O/r02r_ncu_gn_c2 python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > $O/r02r_ncu_gn_c2.log 2>&1; echo "ncu gn rc=$?"
ls -la $O/r02r_*
