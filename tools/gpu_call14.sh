#!/bin/bash
mkdir -p gpurun_out
for sgp in 0 1; do
  echo "== probing sub-group $sgp"
  timeout 200 python tools/gpu_rec_profile.py 512 64 1500 $sgp 2>&1 | tee gpurun_out/recprobe_fwd_pair_sg$sgp.txt
done
