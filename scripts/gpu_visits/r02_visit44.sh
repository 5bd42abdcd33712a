#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_nets.py -q -x -k "bit_equal" 2>&1 | tail -5
