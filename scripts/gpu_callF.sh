#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_train_gpu.py -q -x --timeout 200 -p no:cacheprovider > gpurun_out/cF_pytest.log 2>&1
echo "pytest train rc=$?"; tail -n 6 gpurun_out/cF_pytest.log
for tc in 3 5; do echo "TRAIN_TC=$tc"; ROKO_B200_TRAIN_TC=$tc timeout 200 python scripts/train_profile.py 128 20; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/cF_train_launches.csv python scripts/train_profile.py 128 2 > gpurun_out/cF_train_ncu.log 2>&1
echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"tn_tc_kernel|tc_stream_kernel|embed_drop" --launch-skip 14 -c 7 -f -o gpurun_out/cF_train_full python scripts/train_profile.py 128 1 > gpurun_out/cF_train_full.log 2>&1
echo "ncu full rc=$?"; ls -la gpurun_out/cF_train_full.ncu-rep
