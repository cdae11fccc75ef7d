#!/bin/bash
set -u
mkdir -p gpurun_out
DIAG_SIZES=128,2368 timeout 400 python scripts/gpu_diag.py "CTA-pair" > gpurun_out/c14_diag.log 2>&1
echo "diag rc=$?"; cat gpurun_out/c14_diag.log | cut -c1-1600
