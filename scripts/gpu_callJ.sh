#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py -q -x --timeout 200 -p no:cacheprovider > gpurun_out/cK_pytest.log 2>&1
echo "pytest train rc=$?"; tail -n 5 gpurun_out/cK_pytest.log
timeout 400 python -m pytest tests/test_variants_gpu.py -q -k "TRAIN_TC" --timeout 200 -p no:cacheprovider > gpurun_out/cK_variants.log 2>&1
echo "variants rc=$?"; tail -n 4 gpurun_out/cK_variants.log
for tc in 6; do echo "TRAIN_TC=$tc"; ROKO_B200_TRAIN_TC=$tc timeout 200 python scripts/train_profile.py 128 20; done
ROKO_B200_TRAIN_TC=6 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/cK_train_launches.csv python scripts/train_profile.py 128 2 > gpurun_out/cK_train_ncu.log 2>&1
echo "ncu rc=$?"
