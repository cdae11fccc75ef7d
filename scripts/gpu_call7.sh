#!/bin/bash
set -u
mkdir -p gpurun_out
DIAG_SIZES=128,2368 timeout 600 python scripts/gpu_diag.py "default:" > gpurun_out/c7_diag.log 2>&1
echo "diag rc=$?"; cat gpurun_out/c7_diag.log | cut -c1-1500
timeout 1200 python -m pytest tests/test_parity_gpu.py tests/test_variants_gpu.py -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/c7_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 12 gpurun_out/c7_pytest.log
