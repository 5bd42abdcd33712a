#!/bin/bash
# one-pass (online softmax) fused attention: parity + A/B against the two-pass kernel
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ops.py -q -x -k "attention" > $O/r02zc_pytest_attn.log 2>&1; echo "pytest attn rc=$?"; tail -6 $O/r02zc_pytest_attn.log
timeout 1200 python -m pytest tests/test_gpu_adm.py tests/test_gpu_nets.py tests/test_gpu_edm.py -q -x > $O/r02zc_pytest_nets.log 2>&1; echo "pytest nets rc=$?"; tail -4 $O/r02zc_pytest_nets.log
for m in 1 0; do
NLC_ATTN_ONEPASS=$m timeout 600 python scripts/step_profile.py c2 256 fp16 > $O/r02zc_step_c2_fp16_a$m.log 2>&1; head -3 $O/r02zc_step_c2_fp16_a$m.log | tail -2; grep attention $O/r02zc_step_c2_fp16_a$m.log
NLC_ATTN_ONEPASS=$m timeout 600 python scripts/step_profile.py adm256 32 fp16 > $O/r02zc_step_adm_fp16_a$m.log 2>&1; head -3 $O/r02zc_step_adm_fp16_a$m.log | tail -2; grep attention $O/r02zc_step_adm_fp16_a$m.log
done
