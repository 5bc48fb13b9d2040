#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_blstm_gpu.py tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -3
run() { echo "$1 :: $(env $1 timeout 300 python bench.py --workload c3 --steps 30 --warmup 4 --no-cpu-baseline --no-e2e 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); k=d["kernels"]; print(round(d["ms_per_step"],3), d["clocks"]["sm_mhz"])')"; }
for rep in 1 2; do
run "LCB_L0_RELEASED=0"
run "LCB_L0_RELEASED=1"
done
LCB_L0_RELEASED=1 timeout 300 python tools/gpu_timeline.py > gpurun_out/timeline_c3_l0rel.txt 2>&1
