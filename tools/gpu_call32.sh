#!/bin/bash
# bench lines of the final commit (CPU baseline now at the workload's batch width)
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_c3_1gpu.json 2> gpurun_out/bench_c3_1gpu.err; echo "bench rc=$?"; cut -c1-160 gpurun_out/bench_c3_1gpu.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_c3_reference_arm.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-120 gpurun_out/bench_c3_reference_arm.json
for wl in c1 c2; do timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl rc=$?"; done
timeout 600 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt; tail -2 gpurun_out/pytest_gpu.txt
