#!/bin/bash
# 8-GPU call: bench at N=8 (infer + train = BASELINE configs[4]), config 3/4 through the product entry point (1 M windows)
set -u
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521"
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 600 $TR bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/g8_bench.json 2> gpurun_out/g8_bench.err
echo "bench8 rc=$?"; tail -c 400 gpurun_out/g8_bench.err; python - <<'PY'
import json
for l in open('gpurun_out/g8_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value']), d['ms_per_step'], d['blocks'], round(d['e2e']['value']), d['parity']['multi_gpu'])
PY
timeout 600 $TR bench.py --mode train --gpus 8 --steps 50 --warmup 5 > gpurun_out/g8_train.json 2> gpurun_out/g8_train.err
echo "train8 rc=$?"; grep '^{' gpurun_out/g8_train.json | cut -c1-400
timeout 900 $TR scripts/run_config4.py 1000000 gpurun_out/g8_config4.json > gpurun_out/g8_config4.log 2>&1
echo "config4 rc=$?"; tail -n 3 gpurun_out/g8_config4.log | cut -c1-1200
