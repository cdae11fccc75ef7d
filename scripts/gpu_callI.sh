#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 300 python scripts/gpu_diag.py default 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    if l.startswith('DIAG '):
        d=json.loads(l[5:]); print({k:v for k,v in d.items() if 'stage_ms' in k or 'windows_per_s' in k or 'labels' in k or k.startswith('taps@128')})
    else: print(l.strip()[:300])
"
timeout 1700 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider -x > gpurun_out/cI_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 4 gpurun_out/cI_pytest.log
ROKO_B200_TRAIN_TC=5 timeout 200 python scripts/train_profile.py 128 20
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/cI_train_launches.csv python scripts/train_profile.py 128 2 > gpurun_out/cI_train_ncu.log 2>&1
echo "ncu rc=$?"
