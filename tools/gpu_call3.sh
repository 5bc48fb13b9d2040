#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -15 gpurun_out/pytest_gpu.txt
timeout 300 python tools/gpu_rec_profile.py 512 64 > gpurun_out/recprobe_fwd_v2_l2.txt 2>&1
cat gpurun_out/recprobe_fwd_v2_l2.txt
LCB_REC_BG=16 timeout 300 python tools/gpu_rec_profile.py 512 48 > gpurun_out/recprobe_fwd_v2_l2_bg16_b48.txt 2>&1
cat gpurun_out/recprobe_fwd_v2_l2_bg16_b48.txt
timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_v2l2.json 2> gpurun_out/bench_c3_v2l2.err; echo "bench rc=$?"
cat gpurun_out/bench_c3_v2l2.json
