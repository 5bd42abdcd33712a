#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q -x > $O/r02n_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -8 $O/r02n_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02n_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $O/r02n_smoke.log
timeout 1500 python bench.py > $O/r02n_bench.json 2> $O/r02n_bench.err; echo "bench rc=$?"; tail -c 3000 $O/r02n_bench.json; tail -5 $O/r02n_bench.err
