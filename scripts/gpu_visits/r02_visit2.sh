#!/bin/bash
O=gpurun_out; mkdir -p $O
python scripts/r02_graph_diag.py > $O/r02b_graph_diag.log 2>&1; echo "diag rc=$?"; cat $O/r02b_graph_diag.log
NLC_CTA_PAIRS=0 python scripts/r02_graph_diag.py > $O/r02b_graph_diag_nopairs.log 2>&1; echo "diag(no pairs) rc=$?"; cat $O/r02b_graph_diag_nopairs.log
python bench.py --no-cpu-baseline --no-extras --precision fp16 > $O/r02b_bench_c2_fp16.json 2> $O/r02b_bench_c2_fp16.err; echo "bench rc=$?"; cat $O/r02b_bench_c2_fp16.json; tail -5 $O/r02b_bench_c2_fp16.err
python bench.py --no-cpu-baseline --no-extras --precision fp16 --no-graph > $O/r02b_bench_c2_fp16_nograph.json 2> $O/r02b_bench_c2_fp16_nograph.err; echo "bench rc=$?"; cat $O/r02b_bench_c2_fp16_nograph.json
