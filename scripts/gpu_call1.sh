#!/bin/bash
# round-2 GPU call 1: kernel A/B diagnostic, GPU test suite, bench, memcheck
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_smi.log 2>&1
timeout 900 python scripts/gpu_diag.py > gpurun_out/c1_diag.log 2>&1
echo "diag rc=$?"; tail -n 12 gpurun_out/c1_diag.log | cut -c1-1500
timeout 1700 python -m pytest tests -m gpu -q --timeout 300 -p no:cacheprovider > gpurun_out/c1_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 40 gpurun_out/c1_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err
echo "bench rc=$?"; tail -c 1500 gpurun_out/c1_bench.err; head -c 3000 gpurun_out/c1_bench.json
ROKO_B200_REC_TC_MIN=32 timeout 500 compute-sanitizer --tool memcheck --print-limit 30 python scripts/profile_target.py 40 1 > gpurun_out/c1_memcheck.log 2>&1
echo "memcheck rc=$?"; tail -n 8 gpurun_out/c1_memcheck.log
