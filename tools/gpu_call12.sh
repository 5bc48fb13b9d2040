#!/bin/bash
# A/B: G-prefetch issue delayed behind the m_t exchange (LCB_REC_OPT=1); chain length of 16-utterance groups (B=32 -> BG=16)
mkdir -p gpurun_out
for opt in 0 1; do
  LCB_REC_OPT=$opt timeout 200 python tools/gpu_rec_profile.py 512 64 1500 > gpurun_out/recprobe_fwd_opt$opt.txt 2>&1
  echo "== opt $opt"; head -3 gpurun_out/recprobe_fwd_opt$opt.txt
  LCB_REC_OPT=$opt timeout 200 python tools/gpu_rec_insitu.py 1500 > gpurun_out/rec_insitu_opt$opt.txt 2>&1; cat gpurun_out/rec_insitu_opt$opt.txt
done
timeout 200 python tools/gpu_rec_profile.py 512 32 1500 > gpurun_out/recprobe_fwd_B32.txt 2>&1; cat gpurun_out/recprobe_fwd_B32.txt
timeout 200 python tools/gpu_rec_profile_bwd.py 512 32 1500 > gpurun_out/recprobe_bwd_B32.txt 2>&1; cat gpurun_out/recprobe_bwd_B32.txt
LCB_REC_OPT=1 timeout 300 python -m pytest tests/test_blstm_gpu.py -x -q 2>&1 | tail -2
