#!/bin/bash
# A/B: last cut snapped to whole GEMM waves; output layer's dW / db on the side stream
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_model_gpu.py tests/test_blstm_gpu.py tests/test_graph_api_gpu.py -x -q -m gpu 2>&1 | tail -3
run() { echo "$1 :: $(env $1 timeout 300 python bench.py --workload c3 --steps 30 --warmup 4 --no-cpu-baseline --no-e2e 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); print(round(d["ms_per_step"],3), d["clocks"]["sm_mhz"])')"; }
for rep in 1 2; do
run "LCB_SNAP_CUT=0 LCB_MOS_SIDE=0"
run "LCB_SNAP_CUT=1 LCB_MOS_SIDE=0"
run "LCB_SNAP_CUT=0 LCB_MOS_SIDE=1"
run "LCB_SNAP_CUT=1 LCB_MOS_SIDE=1"
done
