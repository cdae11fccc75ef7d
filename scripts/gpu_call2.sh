#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 400 python scripts/gpu_diag.py "front tcgen05" > gpurun_out/c2_diag.log 2>&1
echo "diag rc=$?"; cat gpurun_out/c2_diag.log | cut -c1-2500
