#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_training.py -q > $O/r02x_pytest.log 2>&1; echo "pytest rc=$?"; grep -n "passed\|failed\|FAILED\|Error" $O/r02x_pytest.log | tail -8
timeout 300 python scripts/train_bench.py 128 > $O/r02x_train_bench.log 2>&1; echo "train bench rc=$?"; tail -4 $O/r02x_train_bench.log
NLC_GRAPH=0 timeout 300 python scripts/train_bench.py 128 > $O/r02x_train_bench_nograph.log 2>&1; tail -2 $O/r02x_train_bench_nograph.log
NLC_TRAIN_TC=0 timeout 300 python scripts/train_bench.py 128 > $O/r02x_train_bench_notc.log 2>&1; tail -2 $O/r02x_train_bench_notc.log
