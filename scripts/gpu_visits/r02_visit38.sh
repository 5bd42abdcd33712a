#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q > $O/r02zj_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02zj_pytest_gpu.log
for m in 1 0; do
NLC_TMA_EPI=$m timeout 600 python scripts/step_profile.py c2 256 fp16 > $O/r02zj_step_c2_fp16_e$m.log 2>&1; head -5 $O/r02zj_step_c2_fp16_e$m.log | tail -4
done
NLC_TMA_EPI=1 timeout 300 python scripts/epi_ablate.py 1 > $O/r02zj_epi_ablate_e1.log 2>&1; grep " op" $O/r02zj_epi_ablate_e1.log | grep -v f32 | head -6
timeout 600 python scripts/repro_check.py bf16 8 2>&1 | tail -6
