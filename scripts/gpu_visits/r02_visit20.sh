#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 2400 python -m pytest tests -m gpu -q -x > $O/r02t_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 $O/r02t_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $O/r02t_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 $O/r02t_smoke.log
timeout 600 python scripts/step_profile.py c2 256 fp16 > $O/r02t_step_c2_fp16.log 2>&1; head -40 $O/r02t_step_c2_fp16.log
