#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/gpu_gemm_small.py > gpurun_out/gemm_shapes.txt 2>&1; tail -10 gpurun_out/gemm_shapes.txt
for hf in 0.3 0.36 0.42 0.5; do
  LCB_HEAD_FRAC=$hf timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_hf$hf.json 2> gpurun_out/bench_hf.err; echo "hf=$hf $(cut -c1-160 gpurun_out/bench_hf$hf.json)"
done
