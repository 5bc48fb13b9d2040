#!/bin/bash
# early-rows backward schedule: parity tests, then C3 step time with the schedule off / on / other fractions and priorities
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_blstm_gpu.py -x -q -m gpu 2>&1 | tail -5
run() { echo "$1 :: $(env $1 timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); k=d["kernels"]; print(round(d["ms_per_step"],3), {n:round(v["ms_total"],2) for n,v in k.items()})')"; }
run "LCB_BWD_EARLY_FRAC=0"
run "LCB_BWD_EARLY_FRAC=0.75"
run "LCB_BWD_EARLY_FRAC=0.75 LCB_XSTREAM_PRIO=-1"
run "LCB_BWD_EARLY_FRAC=0.7"
run "LCB_BWD_EARLY_FRAC=0.8 LCB_EARLY_CAP=64"
run "LCB_BWD_EARLY_FRAC=0.65 LCB_EARLY_CAP=148"
