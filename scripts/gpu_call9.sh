#!/bin/bash
set -u
mkdir -p gpurun_out
for ns in 8 12 16; do
  timeout 300 python bench.py --steps 20 --warmup 5 --streams $ns --no-cpu-baseline --no-train --no-library-baseline > gpurun_out/c9_bench_s$ns.json 2> gpurun_out/c9_bench_s$ns.err
  python -c "
import json; d=json.load(open('gpurun_out/c9_bench_s$ns.json')); print('streams $ns', round(d['value']), d['ms_per_step'], round(d['e2e']['value']), round(d['coalesced']['windows_per_s_per_gpu']))"
done
