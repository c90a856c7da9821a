#!/bin/bash
# BASELINE configs[3] on one box: the row-sharded 736x1280 correlation at 1, 2, 4, 8 GPUs (as many as the box has).
# Usage (GPU box): bash tools/run_config4.sh [max_gpus]   -> gpurun_out/config4_r2.jsonl
MAXG=${1:-$(nvidia-smi -L | wc -l)}
mkdir -p gpurun_out
rm -f gpurun_out/config4_r2.jsonl
for G in 1 2 4 8; do
  if [ "$G" -le "$MAXG" ]; then
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port $((29600 + G)) \
      tools/corr_rowshard_bench.py --out gpurun_out/config4_r2.jsonl 2>&1 | grep -v "^W\|^\*\*\*" | tail -6
  fi
done
