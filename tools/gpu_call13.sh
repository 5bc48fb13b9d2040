#!/bin/bash
# paired 16-utterance sub-groups in the forward recurrence: parity + A/B timing
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_blstm_gpu.py tests/test_full_size_gpu.py tests/test_model_gpu.py -x -q 2>&1 | tail -5
for pair in 0 1; do
  echo "== LCB_REC_PAIR=$pair"
  LCB_REC_PAIR=$pair timeout 200 python tools/gpu_rec_profile.py 512 64 1500 > gpurun_out/recprobe_fwd_pair$pair.txt 2>&1; cat gpurun_out/recprobe_fwd_pair$pair.txt
  LCB_REC_PAIR=$pair timeout 200 python tools/gpu_rec_insitu.py 1500 2>&1 | tee gpurun_out/rec_insitu_pair$pair.txt
done
