#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python scripts/repro_check.py fp16 4 > $O/r02zh_repro_fp16.log 2>&1; echo "rc=$?"; cat $O/r02zh_repro_fp16.log | tail -8
timeout 900 python scripts/repro_check.py bf16 4 > $O/r02zh_repro_bf16.log 2>&1; echo "rc=$?"; cat $O/r02zh_repro_bf16.log | tail -8
