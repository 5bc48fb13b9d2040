#!/bin/bash
# vectorised mixture backward, side-stream conversions, layer-0 dWpT beside its BPTT: parity tests + C3/C2 step times
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_model_gpu.py tests/test_blstm_gpu.py tests/test_full_size_gpu.py -x -q -m gpu 2>&1 | tail -5
run() { echo "$1 $2 :: $(env $1 timeout 300 python bench.py --workload $2 --steps 8 --warmup 3 --no-cpu-baseline --no-e2e 2>/dev/null | python -c 'import json,sys; d=json.loads(sys.stdin.read()); k=d["kernels"]; print(round(d["ms_per_step"],3), {n:round(v["ms_total"],2) for n,v in k.items()}, d["roofline"].get("gemm_classes"))')"; }
run "LCB_X=0" c3
run "LCB_X=0" c3
run "LCB_BWD_EARLY_FRAC=0.72" c3
run "LCB_X=0" c2
run "LCB_X=0" c1
