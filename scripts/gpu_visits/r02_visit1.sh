#!/bin/bash
# GPU visit 1 of round 2: accuracy at the benchmark architectures, CUDA-graph check, round-start bench numbers.
O=gpurun_out; mkdir -p $O
python scripts/r02_report.py c2 graph > $O/r02a_report_c2.log 2>&1; echo "report c2 rc=$?"; cat $O/r02a_report_c2.log
python scripts/r02_report.py nets adm > $O/r02a_report_nets.log 2>&1; echo "report nets rc=$?"; cat $O/r02a_report_nets.log
python bench.py --no-cpu-baseline > $O/r02a_bench_c2.json 2> $O/r02a_bench_c2.err; echo "bench c2 rc=$?"; cat $O/r02a_bench_c2.json; tail -3 $O/r02a_bench_c2.err
python -m pytest tests -m gpu -x -q > $O/r02a_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02a_pytest_gpu.log
