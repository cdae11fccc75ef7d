#!/bin/bash
set -u
mkdir -p gpurun_out
DIAG_SIZES=128,2368 timeout 600 python scripts/gpu_diag.py "W-stationary" > gpurun_out/c11_diag.log 2>&1
echo "diag rc=$?"; cat gpurun_out/c11_diag.log | cut -c1-1500
timeout 900 python -m pytest tests/test_parity_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider -k "alternative or full_pass" > gpurun_out/c11_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 8 gpurun_out/c11_pytest.log
