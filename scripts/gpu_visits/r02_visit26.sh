#!/bin/bash
# TMA epilogue (tensor store of the 16-bit output, tensor load of the 16-bit residual): parity, ablation A/B, c2 step A/B
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ops.py -q -x > $O/r02z_pytest_ops.log 2>&1; echo "pytest ops rc=$?"; tail -15 $O/r02z_pytest_ops.log
NLC_TMA_EPI=1 timeout 300 python scripts/epi_ablate.py 1 > $O/r02z_epi_ablate_tma1.log 2>&1; echo "epi1 rc=$?"; grep "op" $O/r02z_epi_ablate_tma1.log | grep -v f32 | head -20
NLC_TMA_EPI=0 timeout 300 python scripts/epi_ablate.py 1 > $O/r02z_epi_ablate_tma0.log 2>&1; echo "epi0 rc=$?"; grep "op" $O/r02z_epi_ablate_tma0.log | grep -v f32 | head -20
NLC_TMA_EPI=1 timeout 600 python scripts/step_profile.py c2 256 fp16 > $O/r02z_step_c2_fp16_tma1.log 2>&1; head -12 $O/r02z_step_c2_fp16_tma1.log
NLC_TMA_EPI=0 timeout 600 python scripts/step_profile.py c2 256 fp16 > $O/r02z_step_c2_fp16_tma0.log 2>&1; head -12 $O/r02z_step_c2_fp16_tma0.log
