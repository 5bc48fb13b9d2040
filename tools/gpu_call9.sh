#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -8 gpurun_out/pytest_gpu.txt
timeout 300 python tools/gpu_rec_profile_bwd.py 512 64 > gpurun_out/recprobe_bwd_v3b.txt 2>&1
cat gpurun_out/recprobe_bwd_v3b.txt
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c3_v3d.json 2> gpurun_out/bench_c3_v3d.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c3_v3d.json'))
print(d['ms_per_step'], {k:round(v['ms_total'],2) for k,v in d['kernels'].items()}, d['config']['final_loss'], d['config']['device_error'])
PY
