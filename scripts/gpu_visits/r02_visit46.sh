#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python scripts/train_bench.py 128 > $O/r02zo_train_bench.log 2>&1; echo "rc=$?"; tail -5 $O/r02zo_train_bench.log
