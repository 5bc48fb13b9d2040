#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python tools/gpu_timeline.py > gpurun_out/timeline_c3.txt 2>&1; head -60 gpurun_out/timeline_c3.txt
timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; cut -c1-200 gpurun_out/bench_quick.json
for wl in c1 c2; do timeout 300 python bench.py --workload $wl --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; cut -c1-200 gpurun_out/bench_$wl.json; done
