#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt
tail -12 gpurun_out/pytest_gpu.txt
timeout 900 python tools/ctc_sweep.py > gpurun_out/ctc_sweep_b256_split.jsonl 2> gpurun_out/ctc_sweep.err; echo "sweep rc=$?"
python - <<'PY'
import json
old={ (r['T'],r['L'],r['V']): r for r in map(json.loads, open('profiles/r01_ctc_sweep_b256.jsonl')) }
for line in open('gpurun_out/ctc_sweep_b256_split.jsonl'):
    r=json.loads(line); k=(r['T'],r['L'],r['V']); o=old.get(k)
    print("T%-5d L%-4d V%-5d %8.3f ms %10.0f utts/s %7.1f GB/s (%.3f)  was %8.3f ms" % (r['T'],r['L'],r['V'],r['ms'],r['utts_per_s'],r['gbs'],r['frac_of_hbm_peak'], o['ms'] if o else -1))
PY
timeout 600 python bench.py --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_c3_v3c.json 2> gpurun_out/bench_c3_v3c.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_c3_v3c.json'))
print(d['ms_per_step'], {k:round(v['ms_total'],2) for k,v in d['kernels'].items()}, d['config']['final_loss'], d['config']['device_error'])
PY
