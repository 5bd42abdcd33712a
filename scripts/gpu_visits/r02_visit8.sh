#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 300 python scripts/gn_bench.py > $O/r02h_gn_bench.log 2>&1; echo "gn rc=$?"; cat $O/r02h_gn_bench.log
timeout 300 python scripts/power_probe.py > $O/r02h_power_probe.log 2>&1; echo "power rc=$?"; cat $O/r02h_power_probe.log
timeout 600 python scripts/ref_gpu_selfparity.py > $O/r02h_ref_gpu_selfparity.log 2>&1; echo "selfparity rc=$?"; tail -5 $O/r02h_ref_gpu_selfparity.log
