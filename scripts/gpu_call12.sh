#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/c12_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 6 gpurun_out/c12_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/c12_smoke.log 2>&1
echo "smoke rc=$?"; tail -n 3 gpurun_out/c12_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c12_bench.json 2> gpurun_out/c12_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/c12_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/c12_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','blocks','timed_region_s')}, round(d['e2e']['value']), d['clocks'])
print(d['roofline']); print(d['coalesced']['windows_per_s_per_gpu'], d['coalesced']['kernels']); print(d['vs_library']); print(d.get('training')); print(d.get('cpu_baseline'))
PY
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/c12_bench_ref.json 2>/dev/null; head -c 500 gpurun_out/c12_bench_ref.json
timeout 120 scripts/ubench/rec_trace
