#!/bin/bash
O=gpurun_out; mkdir -p $O
for mask in 9 10 11 15; do
echo "mask $mask"
NLC_TMA_EPI_MASK=$mask timeout 600 python scripts/repro_check.py fp16 16 2>&1 | tail -1
done
timeout 600 compute-sanitizer --tool racecheck --racecheck-report all python -m pytest tests/test_gpu_ops.py -q -x -k "tma_epilogue and fp16 and (case0 or case3 or case5)" > $O/r02zi_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -15 $O/r02zi_racecheck.log
