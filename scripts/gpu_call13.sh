#!/bin/bash
set -u
mkdir -p gpurun_out
for mode in 1 0; do timeout 20 scripts/ubench/cta2_gemm $mode; echo "mode $mode rc=$?"; done
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-train --no-library-baseline > gpurun_out/c13_bench.json 2> gpurun_out/c13_bench.err
echo "bench rc=$?"; tail -c 300 gpurun_out/c13_bench.err; python -c "
import json; d=json.load(open('gpurun_out/c13_bench.json')); print({k:d[k] for k in ('value','ms_per_step','blocks','timed_region_s')}, round(d['e2e']['value']), d['e2e']['timed_region_s'], d['clocks'])"
