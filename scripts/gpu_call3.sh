#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 600 python scripts/gpu_diag.py reference "front mma.sync" default > gpurun_out/c3_diag.log 2>&1
echo "diag rc=$?"; cat gpurun_out/c3_diag.log | cut -c1-1800
timeout 1700 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/c3_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 30 gpurun_out/c3_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c3_bench.json 2> gpurun_out/c3_bench.err
echo "bench rc=$?"; tail -c 1000 gpurun_out/c3_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/c3_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','blocks','timed_region_s')}, d['e2e']['value'], d['clocks'])
print(d['stage_ms']); print(d['coalesced']['stage_ms'], d['coalesced']['windows_per_s_per_gpu']); print(d['vs_library']); print(d.get('training')); print(d.get('cpu_baseline'))
PY
