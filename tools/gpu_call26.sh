#!/bin/bash
# multi-cut early rows + forward head fractions: parity tests, A/B of the C3 step time in one call (same box), timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_blstm_gpu.py -x -q -m gpu 2>&1 | tail -3
run() { echo "$1 :: $(env $1 timeout 300 python bench.py --workload c3 --steps 8 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); k=d["kernels"]; print(round(d["ms_per_step"],3), {n:round(v["ms_total"],2) for n,v in k.items()})')"; }
run "LCB_BWD_EARLY_FRACS="
run "LCB_BWD_EARLY_FRACS=0.7"
run "LCB_BWD_EARLY_FRACS=0.67,0.85"
run "LCB_BWD_EARLY_FRACS=0.7,0.88"
run "LCB_BWD_EARLY_FRACS=0.65,0.8,0.92"
run "LCB_BWD_EARLY_FRACS=0.7 LCB_HEAD_FRACS=0.2,0.5"
run "LCB_BWD_EARLY_FRACS=0.7 LCB_HEAD_FRACS=0.25,0.6"
run "LCB_BWD_EARLY_FRACS=0.7 LCB_HEAD_FRACS=0.3"
run "LCB_BWD_EARLY_FRACS=0.7"
LCB_BWD_EARLY_FRACS=0.67,0.85 timeout 300 python tools/gpu_timeline.py > gpurun_out/timeline_c3_cuts2.txt 2>&1
