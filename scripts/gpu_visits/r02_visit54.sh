#!/bin/bash
# final-binary precision record: the c2 100-step loop gates (printed PSNRs) and the four-mode precision table
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_bench_arch.py -q -s -k "c2" > $O/r02zw_bench_arch_c2.log 2>&1; echo "rc=$?"; grep -n "dB\|passed\|failed" $O/r02zw_bench_arch_c2.log | head -20
timeout 900 python scripts/precision_report.py > $O/r02zw_precision_report.log 2>&1; echo "rc=$?"; tail -12 $O/r02zw_precision_report.log
