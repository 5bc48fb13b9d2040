#!/bin/bash
# Quick GPU visit: parity tests + one c3 bench line without the CPU baseline.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -4 gpurun_out/pytest_gpu.txt
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline "$@" > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print(d['ms_per_step'], d['value'], 'e2e', d['e2e'] and d['e2e']['ms_per_step'], {k:(v['launches'],round(v['ms_total'],2)) for k,v in d['kernels'].items()}, d['config']['final_loss'], d['config']['device_error'], d['gpu_launches'])
PY
