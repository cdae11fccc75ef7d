#!/bin/bash
set -u
mkdir -p gpurun_out
T="timeout 400"
$T ncu --set full --clock-control none --import-source on -k regex:'front_tc_kernel|proj_h_kernel|rec_h_kernel' -s 8 -c 7 \
    -o gpurun_out/r2a_b2368 python scripts/profile_target.py 2368 2 > gpurun_out/c4_ncu.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/c4_ncu.log
ls -la gpurun_out/*.ncu-rep | tail -3
