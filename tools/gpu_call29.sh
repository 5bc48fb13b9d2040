#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_c3_1gpu.json 2> gpurun_out/bench_c3_1gpu.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/bench_c3_1gpu.json
for wl in c1 c2; do timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl rc=$?"; done
