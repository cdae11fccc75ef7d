#!/bin/bash
# 2-GPU call: DDP / sharded-inference tests, bench at N=2 (infer + train), config 4 through the product entry point
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ddp_gpu.py -m gpu -q --timeout 400 -p no:cacheprovider > gpurun_out/c5_pytest.log 2>&1
echo "pytest rc=$?"; tail -n 8 gpurun_out/c5_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c5_bench2.json 2> gpurun_out/c5_bench2.err
echo "bench2 rc=$?"; tail -c 600 gpurun_out/c5_bench2.err; python -c "
import json; d=json.load(open('gpurun_out/c5_bench2.json')); print(d['value'], d['ms_per_step'], d['blocks'], d['e2e']['value'], d['parity'])"
timeout 600 $TR bench.py --mode train --gpus 2 --steps 30 --warmup 5 > gpurun_out/c5_train2.json 2> gpurun_out/c5_train2.err
echo "train2 rc=$?"; tail -c 300 gpurun_out/c5_train2.err; head -c 600 gpurun_out/c5_train2.json; echo
timeout 900 $TR scripts/run_config4.py 200000 gpurun_out/c5_config4_2gpu.json > gpurun_out/c5_config4.log 2>&1
echo "config4 rc=$?"; tail -n 4 gpurun_out/c5_config4.log | cut -c1-900
