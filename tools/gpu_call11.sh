#!/bin/bash
# Session-5 first visit: does the restored tree still pass on the box; recurrence cost at full sequence length.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -3 gpurun_out/pytest_gpu.txt
timeout 200 python tools/gpu_rec_insitu.py 1500 > gpurun_out/rec_insitu.txt 2>&1; cat gpurun_out/rec_insitu.txt
timeout 200 python tools/gpu_rec_profile.py 512 64 1500 > gpurun_out/recprobe_fwd_T1500.txt 2>&1
timeout 200 python tools/gpu_rec_profile_bwd.py 512 64 1500 > gpurun_out/recprobe_bwd_T1500.txt 2>&1
timeout 200 python tools/gpu_rec_profile.py 512 64 200 > gpurun_out/recprobe_fwd_T200.txt 2>&1
timeout 200 python tools/gpu_rec_profile_bwd.py 512 64 200 > gpurun_out/recprobe_bwd_T200.txt 2>&1
timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
cat gpurun_out/bench_quick.json | cut -c1-600
