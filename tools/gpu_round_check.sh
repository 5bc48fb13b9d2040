#!/bin/bash
# One GPU-box visit: parity tests, probes, bench, ncu launch list, one full capture of the projection GEMM.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -3 gpurun_out/pytest_gpu.txt
timeout 300 python tools/gpu_rec_profile.py 512 64 > gpurun_out/recprobe_fwd.txt 2>&1
timeout 300 python tools/gpu_rec_profile_bwd.py 512 64 > gpurun_out/recprobe_bwd.txt 2>&1
timeout 900 python bench.py --steps 6 --warmup 3 > gpurun_out/bench_c3.json 2> gpurun_out/bench_c3.err; echo "bench rc=$?"
cat gpurun_out/bench_c3.json
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_bf16 -s 60 -c 3 -o gpurun_out/prof_gemm -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_gemm.log 2>&1; echo "ncu full rc=$?"
