#!/bin/bash
set -u
mkdir -p gpurun_out
bash scripts/capture_profiles.sh > gpurun_out/cL_capture.log 2>&1
tail -n 12 gpurun_out/cL_capture.log
timeout 1700 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/cL_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 4 gpurun_out/cL_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/cL_smoke.log 2>&1
echo "smoke rc=$?"; tail -n 2 gpurun_out/cL_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/cL_bench.json 2> gpurun_out/cL_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/cL_bench.err; python - <<'PY'
import json
d=json.load(open('gpurun_out/cL_bench.json'))
print({k:d[k] for k in ('value','ms_per_step','blocks','timed_region_s')}, round(d['e2e']['value']), d['clocks'])
print(d['roofline']['kernel'], d['roofline']['frac'], d['roofline']['traffic']); print(d['stage_ms']); print(d['coalesced']['stage_ms'], d['coalesced']['windows_per_s_per_gpu']); print(d['vs_library']['ratio'], d.get('training',{}).get('ms_per_step'), d.get('cpu_baseline',{}).get('value'))
PY
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/cL_bench_ref.json 2>/dev/null; head -c 300 gpurun_out/cL_bench_ref.json; echo
timeout 120 python bench.py --impl torch_gpu > gpurun_out/cL_bench_torch_gpu.json 2>/dev/null; head -c 300 gpurun_out/cL_bench_torch_gpu.json; echo
