#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 60 scripts/ubench/rec_trace
DIAG_SIZES=128,2368 timeout 400 python scripts/gpu_diag.py "default:" > gpurun_out/cC_diag.log 2>&1
echo "diag rc=$?"; cat gpurun_out/cC_diag.log | cut -c1-1300
