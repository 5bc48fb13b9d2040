#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py > gpurun_out/bench_c3_1gpu.json 2> gpurun_out/bench_c3_1gpu.err; cut -c1-160 gpurun_out/bench_c3_1gpu.json
for cfgs in "0.36 1" "0.3 0" "0.36 0" "0.3 1" "0 0"; do set -- $cfgs
  echo "c2 fracs=$1 hproj=$2 $(LCB_HEAD_FRACS=$1 LCB_OVERLAP_HPROJ=$2 timeout 300 python bench.py --workload c2 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | cut -c1-140)"
done
for cfgs in "0.36 1" "0.3 0"; do set -- $cfgs
  echo "c1 fracs=$1 hproj=$2 $(LCB_HEAD_FRACS=$1 LCB_OVERLAP_HPROJ=$2 timeout 300 python bench.py --workload c1 --steps 10 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | cut -c1-140)"
done
