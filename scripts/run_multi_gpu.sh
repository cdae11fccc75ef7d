#!/bin/bash
# N-GPU runs of one box (under gpurun --gpus N): bench at N (inference + training step), BASELINE configs[3] (1 M synthetic windows
# sharded through the product entry point roko_b200.inference.infer_fast) and configs[4] (epochs of roko_b200.train on 100 k synthetic
# labelled windows, 128 windows per GPU and step).  Usage: bash scripts/run_multi_gpu.sh [N=8] [all|train]
set -u
N=${1:-8}
WHAT=${2:-all}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521"
if [ "$WHAT" = all ]; then
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/g${N}_bench.json 2> gpurun_out/g${N}_bench.err
echo "bench rc=$?"
fi
timeout 600 $TR bench.py --mode train --gpus $N --steps 50 --warmup 5 > gpurun_out/g${N}_train.json 2> gpurun_out/g${N}_train.err
echo "train bench rc=$?"
timeout 900 $TR -m roko_b200.train "synthetic://100000?labels=1" gpurun_out/g${N}_ckpt --val "synthetic://2048?labels=1&seed=9" --memory --b 128 --epochs 3 > gpurun_out/g${N}_config5.log 2>&1
echo "config 5 rc=$?"; grep -E "^Epoch|Val" gpurun_out/g${N}_config5.log | sort | uniq -c | tail -8; rm -rf gpurun_out/g${N}_ckpt
[ "$WHAT" = all ] || exit 0
timeout 900 $TR scripts/run_config4.py 1000000 gpurun_out/g${N}_config4.json > gpurun_out/g${N}_config4.log 2>&1
echo "config 4 rc=$?"; tail -n 1 gpurun_out/g${N}_config4.log | cut -c1-1000
