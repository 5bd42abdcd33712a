#!/bin/bash
# deterministic split-K of the small-M conv launches: parity + A/B at batch 256 and 32
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ops.py -q -x > $O/r02zr_pytest_ops.log 2>&1; echo "pytest ops rc=$?"; tail -4 $O/r02zr_pytest_ops.log
timeout 900 python -m pytest tests/test_gpu_nets.py tests/test_gpu_adm.py tests/test_gpu_edm.py tests/test_gpu_fid.py tests/test_gpu_full_size.py -q -x > $O/r02zr_pytest_nets.log 2>&1; echo "pytest nets rc=$?"; tail -3 $O/r02zr_pytest_nets.log
for b in 256 32; do
for m in 1 0; do
NLC_SPLITK=$m timeout 600 python scripts/step_profile.py c2 $b fp16 10 > $O/r02zr_step_c2_b${b}_k$m.log 2>&1; echo "b$b splitk=$m: $(sed -n 2,2p $O/r02zr_step_c2_b${b}_k$m.log)"; grep "4x4 K4608 N512\|2x2 K4608\|8x8 K2304 N256 s1" $O/r02zr_step_c2_b${b}_k$m.log | head -3
done
done
for m in 1 0; do
NLC_SPLITK=$m timeout 600 python scripts/step_profile.py adm256 16 fp16 > $O/r02zr_step_adm_k$m.log 2>&1; echo "adm b16 splitk=$m: $(sed -n 2,2p $O/r02zr_step_adm_k$m.log)"
done
