#!/bin/bash
# N=2 sweep of gradient-exchange knobs (bucket size, NCCL channel count); prints ms_per_step per setting
run() {
  tag=$1; shift
  env "$@" timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 \
      bench.py --gpus 2 --steps 10 --warmup 3 --no-e2e --no-roofline --no-cpu-baseline --bucket-mb $BMB > gpurun_out/n2_$tag.json 2> gpurun_out/n2_$tag.err
  python -c "import json,sys; d=json.load(open('gpurun_out/n2_$tag.json')); print('$tag', d['ms_per_step'], d['value'])"
}
BMB=32 run base A=1
BMB=32 run ch4 NCCL_MAX_NCHANNELS=4
BMB=32 run ch2 NCCL_MAX_NCHANNELS=2
BMB=2048 run end A=1
BMB=2048 run end_ch8 NCCL_MAX_NCHANNELS=8
BMB=128 run b128_ch4 NCCL_MAX_NCHANNELS=4
