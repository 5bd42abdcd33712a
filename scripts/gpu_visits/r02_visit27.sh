#!/bin/bash
# 16-bit epilogue variants A/B: 0 = staged LSU, 1 = TMA load/store, 2 = 256-bit global accesses from registers
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ops.py -q -x > $O/r02za_pytest_ops.log 2>&1; echo "pytest ops rc=$?"; tail -5 $O/r02za_pytest_ops.log
for m in 0 1 2; do
NLC_TMA_EPI=$m timeout 300 python scripts/epi_ablate.py 1 > $O/r02za_epi_ablate_e$m.log 2>&1; echo "epi$m rc=$?"; grep " op" $O/r02za_epi_ablate_e$m.log | grep -v f32 | head -20
done
for m in 0 1 2; do
NLC_TMA_EPI=$m timeout 600 python scripts/step_profile.py c2 256 fp16 > $O/r02za_step_c2_fp16_e$m.log 2>&1; head -5 $O/r02za_step_c2_fp16_e$m.log
done
for m in 0 1 2; do
NLC_TMA_EPI=$m timeout 600 python scripts/step_profile.py adm256 16 fp16 > $O/r02za_step_adm_fp16_e$m.log 2>&1; head -5 $O/r02za_step_adm_fp16_e$m.log
done
