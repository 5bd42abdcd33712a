#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 600 python scripts/step_profile.py c2 32 fp16 10 > $O/r02zq_step_c2_b32_fp16.log 2>&1; head -45 $O/r02zq_step_c2_b32_fp16.log
