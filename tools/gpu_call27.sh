#!/bin/bash
# A/B of schedule options on one box, 30 timed steps each, interleaved twice
mkdir -p gpurun_out
run() { echo "$1 :: $(env $1 timeout 300 python bench.py --workload c3 --steps 30 --warmup 4 --no-cpu-baseline --no-e2e 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); k=d["kernels"]; print(round(d["ms_per_step"],3), d["clocks"])')"; }
for rep in 1 2; do
run "LCB_BWD_EARLY_FRACS="
run "LCB_BWD_EARLY_FRACS=0.7"
run "LCB_BWD_EARLY_FRACS=0.67,0.85"
run "LCB_BWD_EARLY_FRACS=0.67,0.85 LCB_HEAD_FRACS=0.2,0.5"
run "LCB_BWD_EARLY_FRACS=0.67,0.85 LCB_HEAD_FRACS=0.25,0.6"
run "LCB_BWD_EARLY_FRACS=0.65,0.8,0.92"
done
