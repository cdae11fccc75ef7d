#!/bin/bash
set -u
mkdir -p gpurun_out
ROKO_B200_PROJ=fp16x2 timeout 400 ncu --set full --clock-control none --import-source on -k regex:'proj_h2_kernel' -s 3 -c 2 \
    -o gpurun_out/r2b_projh2 python scripts/profile_target.py 2368 2 > gpurun_out/c15_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/c15_ncu.log
