#!/bin/bash
mkdir -p gpurun_out
for opt in 0 1 2 3; do
  echo "== LCB_REC_OPT=$opt"
  LCB_REC_OPT=$opt timeout 200 python tools/gpu_rec_profile.py 512 64 1500 2>&1 | grep "step period"
  LCB_REC_OPT=$opt timeout 200 python tools/gpu_rec_profile_bwd.py 512 64 1500 2>&1 | grep -v "^BPTT"
  LCB_REC_OPT=$opt timeout 200 python tools/gpu_rec_insitu.py 1500 2>&1 | grep -v "head_frac 0.0"
done
LCB_REC_OPT=2 timeout 300 python -m pytest tests/test_blstm_gpu.py tests/test_model_gpu.py -x -q 2>&1 | tail -2
