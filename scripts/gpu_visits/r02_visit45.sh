#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_training.py -q -x > $O/r02zn_pytest_training.log 2>&1; echo "rc=$?"; tail -15 $O/r02zn_pytest_training.log
