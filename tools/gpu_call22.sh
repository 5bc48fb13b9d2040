#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
run() { echo "fracs=$1 hproj=$2 $(LCB_HEAD_FRACS=$1 LCB_OVERLAP_HPROJ=$2 timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e 2>gpurun_out/bench_hf.err | cut -c1-150)"; }
run 0.36 0
run 0.36 1
run 0.15,0.45 1
run 0.09,0.25,0.55 1
run 0.12,0.33,0.62 1
run 0.09,0.25,0.55 0
