#!/bin/bash
O=gpurun_out; mkdir -p $O
for mask in 10 6 15; do
echo "mask $mask"
NLC_TMA_EPI_MASK=$mask timeout 900 python scripts/repro_check.py fp16 48 2>&1 | tail -1
done
timeout 600 python -m pytest tests/test_gpu_ops.py -q -x -k "tma_epilogue or slab" 2>&1 | tail -2
