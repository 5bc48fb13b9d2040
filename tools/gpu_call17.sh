#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_blstm_gpu.py tests/test_full_size_gpu.py -x -q 2>&1 | tail -2
for pair in 0 1; do
  echo "== PAIR=$pair"
  LCB_REC_PAIR=$pair timeout 200 python tools/gpu_rec_profile.py 512 64 1500 2>&1 | grep -v "^ctl:start\|global_stores\|cluster:"
done
timeout 200 python tools/gpu_rec_profile.py 512 32 1500 2>&1 | grep -v "^ctl:start\|global_stores\|cluster:"
