#!/bin/bash
set -u
mkdir -p gpurun_out
DIAG_SIZES=128,2368,4736 timeout 900 python scripts/gpu_diag.py "default:" pingpong > gpurun_out/c6_diag.log 2>&1
echo "diag rc=$?"; cat gpurun_out/c6_diag.log | cut -c1-2600
timeout 400 ncu --set full --clock-control none --import-source on -k regex:'front_tc_kernel' -s 1 -c 1 \
    -o gpurun_out/r2a_front_b2368 python scripts/profile_target.py 2368 2 > gpurun_out/c6_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/c6_ncu.log
