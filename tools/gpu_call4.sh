#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu_rec_profile.py 512 64 > gpurun_out/recprobe_fwd_v2_l2.txt 2>&1
cat gpurun_out/recprobe_fwd_v2_l2.txt
LCB_REC_XCH=dsmem timeout 300 python tools/gpu_rec_profile.py 512 64 > gpurun_out/recprobe_fwd_v2_dsmem.txt 2>&1
cat gpurun_out/recprobe_fwd_v2_dsmem.txt
LCB_REC_BG=16 timeout 300 python tools/gpu_rec_profile.py 512 16 > gpurun_out/recprobe_fwd_v2_l2_bg16_b16.txt 2>&1
cat gpurun_out/recprobe_fwd_v2_l2_bg16_b16.txt
