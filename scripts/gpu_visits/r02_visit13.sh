#!/bin/bash
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_fid.py -q -x > $O/r02m_pytest_fid.log 2>&1; echo "pytest fid rc=$?"; tail -30 $O/r02m_pytest_fid.log
timeout 900 python -m pytest tests/test_gpu_ops.py tests/test_gpu_nets.py tests/test_gpu_fp32_mode.py -q -x > $O/r02m_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 $O/r02m_pytest.log
timeout 300 python scripts/epi_ablate.py 1 2>&1 | head -9
