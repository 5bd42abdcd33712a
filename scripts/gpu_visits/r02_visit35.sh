#!/bin/bash
O=gpurun_out; mkdir -p $O
for mask in 13 14 7 11 5 6 9 10; do
echo "mask $mask (1 no-resid launches, 2 resid launches, 4 slab kernel, 8 tap kernel)"
NLC_TMA_EPI_MASK=$mask timeout 600 python scripts/repro_check.py fp16 4 2>&1 | tail -1
done
