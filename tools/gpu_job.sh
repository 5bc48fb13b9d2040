mkdir -p gpurun_out
timeout 1700 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r02_pytest_gpu.txt; cat gpurun_out/r02_pytest_gpu.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_c3_1gpu.json 2> gpurun_out/r02_bench_c3_1gpu.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c3_1gpu.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['clocks'], {k:(v['ms_total']) for k,v in d['kernels'].items()}); print(d['roofline']['us_per_time_step'], d['config'])"; tail -3 gpurun_out/r02_bench_c3_1gpu.err
python __graft_entry__.py 2>&1 | tail -2; python -c "
import __graft_entry__ as g; g.smoke(); print('smoke ok')"
