#!/bin/bash
# 8-GPU sweep of the gradient all-reduce settings (wire dtype, NCCL CTA cap, bucket size); one JSON line per setting.
# Usage (on the GPU box, 8 GPUs):  bash tools/n8_sweep.sh [N] > gpurun_out/n8_sweep.jsonl
N=${1:-8}
run() {
  local tag="$1"; shift
  local line
  line=$(env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 \
      bench.py --gpus $N --steps 10 --warmup 3 --no-gpu-eager --no-cpu-baseline --no-e2e $EXTRA 2>gpurun_out/n8_${tag}.err | tail -1)
  echo "{\"tag\": \"$tag\", \"line\": $line}"
}
EXTRA="" run fp32_default X=1
EXTRA="--grad-comm bf16" run bf16_wire X=1
EXTRA="" run fp32_maxctas8 NCCL_MAX_CTAS=8
EXTRA="--grad-comm bf16" run bf16_maxctas8 NCCL_MAX_CTAS=8
EXTRA="--grad-comm bf16 --bucket-mb 64" run bf16_bucket64 X=1
