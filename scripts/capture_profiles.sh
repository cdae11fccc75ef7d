#!/bin/bash
# Run under gpurun (1 GPU): ncu launch list of the bench command, one full capture per hot kernel at the bench batch (128)
# and at one coalesced pass (2368), compute-sanitizer memcheck + racecheck at small batches.  Output: gpurun_out/r02_*.
set -u
mkdir -p gpurun_out
T="timeout 400"
$T ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/r02_launches.csv \
    python bench.py --steps 6 --warmup 3 --min-region 0.02 --no-cpu-baseline --no-train --no-library-baseline > gpurun_out/r02_bench_under_ncu.log 2>&1
echo "launch list rc=$?"
K='front_tc_kernel|proj_h_kernel|rec_h_kernel|head_kernel|rec_kernel'
$T ncu --set full --clock-control none --import-source on -k regex:"$K" -s 8 -c 8 -o gpurun_out/r02_b128 \
    python scripts/profile_target.py 128 3 > /dev/null 2>&1
echo "ncu b128 rc=$?"
$T ncu --set full --clock-control none --import-source on -k regex:"$K" -s 8 -c 8 -o gpurun_out/r02_b2368 \
    python scripts/profile_target.py 2368 3 > /dev/null 2>&1
echo "ncu b2368 rc=$?"
# sanitizers: ragged batch (40 windows -> a partial 32-window group in rec_h) and a sub-threshold batch (FFMA recurrence)
ROKO_B200_REC_TC_MIN=32 ROKO_B200_GRAPHS=0 $T compute-sanitizer --tool memcheck --print-limit 20 python scripts/profile_target.py 40 1 > gpurun_out/r02_memcheck_b40.log 2>&1
echo "memcheck rc=$?"; tail -n 2 gpurun_out/r02_memcheck_b40.log
ROKO_B200_GRAPHS=0 $T compute-sanitizer --tool memcheck --print-limit 20 python scripts/profile_target.py 5 1 > gpurun_out/r02_memcheck_b5.log 2>&1
echo "memcheck b5 rc=$?"; tail -n 2 gpurun_out/r02_memcheck_b5.log
ROKO_B200_REC_TC_MIN=32 ROKO_B200_GRAPHS=0 timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python scripts/profile_target.py 33 1 > gpurun_out/r02_racecheck_b33.log 2>&1
echo "racecheck rc=$?"; tail -n 3 gpurun_out/r02_racecheck_b33.log
# training step: launch list, event timings, sanitizers at a ragged small batch (3 windows = 105.5 row tiles)
$T ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_train_launches.csv \
    python scripts/train_profile.py 128 2 > /dev/null 2>&1
echo "train launch list rc=$?"
$T python scripts/train_profile.py 128 20 > gpurun_out/r02_train_profile.log 2>&1; tail -n 1 gpurun_out/r02_train_profile.log
$T compute-sanitizer --tool memcheck --print-limit 20 python scripts/train_profile.py 3 1 > gpurun_out/r02_memcheck_train_b3.log 2>&1
echo "memcheck train rc=$?"; tail -n 2 gpurun_out/r02_memcheck_train_b3.log
timeout 900 compute-sanitizer --tool racecheck --print-limit 20 python scripts/train_profile.py 2 1 > gpurun_out/r02_racecheck_train_b2.log 2>&1
echo "racecheck train rc=$?"; tail -n 3 gpurun_out/r02_racecheck_train_b2.log
ls -la gpurun_out | grep r02_
