#!/bin/bash
set -u
mkdir -p gpurun_out
DIAG_SIZES=128,2368 timeout 400 python scripts/gpu_diag.py "default:" > gpurun_out/cC_diag.log 2>&1
echo "diag rc=$?"; cat gpurun_out/cC_diag.log | cut -c1-1300
timeout 600 python -m pytest tests/test_parity_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/cC_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 5 gpurun_out/cC_pytest.log
