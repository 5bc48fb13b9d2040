mkdir -p gpurun_out
python tools/bench_infer.py > gpurun_out/r02_bench_c5_infer.json 2> gpurun_out/r02_bench_c5_infer.err; tail -c 1500 gpurun_out/r02_bench_c5_infer.json; tail -3 gpurun_out/r02_bench_c5_infer.err
for wl in c1 c2; do python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-ctc > gpurun_out/r02_bench_$wl.json 2> gpurun_out/r02_bench_$wl.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_$wl.json')); print('$wl', d['ms_per_step'], d['value'], d['e2e']['value'])"; done
