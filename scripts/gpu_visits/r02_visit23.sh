#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_ops.py tests/test_gpu_nets.py tests/test_gpu_adm.py tests/test_gpu_edm.py tests/test_gpu_training.py -q > $O/r02w_pytest.log 2>&1; echo "pytest rc=$?"; grep -n "passed\|failed\|FAILED" $O/r02w_pytest.log | tail -5
timeout 300 python scripts/train_bench.py 128 > $O/r02w_train_bench.log 2>&1; echo "train bench rc=$?"; cat $O/r02w_train_bench.log | tail -4
timeout 600 python scripts/step_profile.py c2 256 fp16 > $O/r02w_step_c2_fp16.log 2>&1; head -3 $O/r02w_step_c2_fp16.log; grep "attention\|4x4 C512\|8x8 C256 " $O/r02w_step_c2_fp16.log | head
