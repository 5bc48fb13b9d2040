#!/bin/bash
# recurrence iteration: parity tests of the encoder/model, clock probes, quick bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_blstm_gpu.py tests/test_model_gpu.py -x -q 2>&1 | tail -3
timeout 300 python tools/gpu_rec_profile.py 512 64 > gpurun_out/recprobe_fwd.txt 2>&1; cat gpurun_out/recprobe_fwd.txt
timeout 300 python tools/gpu_rec_profile_bwd.py 512 64 > gpurun_out/recprobe_bwd.txt 2>&1; cat gpurun_out/recprobe_bwd.txt
timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_quick.json'))
print(d['ms_per_step'], d['value'], {k:(v['launches'],round(v['ms_total'],2)) for k,v in d['kernels'].items()}, d['config']['final_loss'], d['config']['device_error'], d['gpu_launches'])
PY
