#!/bin/bash
# Run under gpurun: ncu launch list of the bench command + one full capture per hot kernel.
set -u
mkdir -p gpurun_out
T="timeout 300"
$T ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/final_launches.csv \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/final_bench_under_ncu.log 2>&1
for spec in "front_kernel 128 1" "proj_tc3_kernel 128 3" "rec_kernel 128 3" "head_kernel 128 1" \
            "front_kernel 2368 1" "proj_tc3_kernel 2368 3" "rec_tc_kernel 2368 3"; do
    set -- $spec
    $T ncu --set full --clock-control none --import-source on -k regex:$1 -s $3 -c 2 \
        -o gpurun_out/final_${1}_b$2 python scripts/profile_target.py $2 3 > /dev/null 2>&1
done
# the tensor-core recurrence at one 128-window batch (the bench's throughput configuration)
ROKO_B200_REC_TC_MIN=128 $T ncu --set full --clock-control none --import-source on -k regex:rec_tc_kernel -s 3 -c 2 \
    -o gpurun_out/final_rec_tc_kernel_b128 python scripts/profile_target.py 128 3 > /dev/null 2>&1
ls -la gpurun_out | grep final_
