#!/bin/bash
# data-parallel overhead at 2 GPUs vs the number of NCCL channels (each channel is a CTA that competes with the recurrence clusters and capped GEMMs)
mkdir -p gpurun_out
run() { echo "$1 :: $(env $1 timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 4 --no-e2e 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],3), round(d["value"]))')"; }
run "LCB_X=0"
run "NCCL_MAX_NCHANNELS=2"
run "NCCL_MAX_NCHANNELS=4"
run "NCCL_MAX_NCHANNELS=8"
run "LCB_X=0"
run "NCCL_MAX_NCHANNELS=2"
