#!/bin/bash
mkdir -p gpurun_out
for pair in 0 1; do for opt in 0 1; do
  echo "== PAIR=$pair OPT=$opt"
  LCB_REC_OPT=$opt LCB_REC_PAIR=$pair timeout 200 python tools/gpu_rec_profile.py 512 64 1500 2>&1 | grep -v "^ctl:start\|global_stores\|cluster:" 
done; done
LCB_REC_OPT=1 timeout 300 python -m pytest tests/test_blstm_gpu.py -x -q 2>&1 | tail -2
