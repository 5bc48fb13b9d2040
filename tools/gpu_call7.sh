#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -8 gpurun_out/pytest_gpu.txt
timeout 900 python bench.py --steps 4 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c3_v3b.json 2> gpurun_out/bench_c3_v3b.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c3_v3b.json'))
print(d['ms_per_step'], d['e2e']['ms_per_step'], {k:round(v['ms_total'],2) for k,v in d['kernels'].items()}, d['config']['final_loss'], d['config']['device_error'])
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_v3b.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
