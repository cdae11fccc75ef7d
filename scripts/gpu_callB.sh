#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python scripts/train_profile.py 128 10
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_train_launches.csv python scripts/train_profile.py 128 2 > /dev/null 2>&1
echo "ncu rc=$?"; python scripts/summarize_launches.py gpurun_out/r02_train_launches.csv embed_drop_kernel | head -40
