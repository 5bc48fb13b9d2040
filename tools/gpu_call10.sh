#!/bin/bash
# Re-entry check of HEAD: parity tests, recurrence probes, bench c3 (full line), ncu launch list.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -5 gpurun_out/pytest_gpu.txt
timeout 300 python tools/gpu_rec_profile.py 512 64 > gpurun_out/recprobe_fwd.txt 2>&1
timeout 300 python tools/gpu_rec_profile_bwd.py 512 64 > gpurun_out/recprobe_bwd.txt 2>&1
tail -15 gpurun_out/recprobe_fwd.txt; tail -15 gpurun_out/recprobe_bwd.txt
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench rc=$?"
cat gpurun_out/bench_c3.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
