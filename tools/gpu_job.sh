# Evidence pass of a GPU visit (run as: gpurun --timeout 3000 -- 'bash tools/gpu_job.sh'; 2 / 4-GPU lines: tools/gpu_job_multi.sh under
# gpurun --gpus N).  Outputs land in gpurun_out/; what is quoted in DESIGN.md is copied to profiles/ (index: profiles/README.md).
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r02_pytest_gpu.txt; cat gpurun_out/r02_pytest_gpu.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_c3_1gpu.json 2> gpurun_out/r02_bench_c3_1gpu.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c3_1gpu.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['clocks'], {k:(v['ms_total']) for k,v in d['kernels'].items()}); print(d['roofline']['us_per_time_step'], d['roofline']['frac'], d['rooflines_other']['gemm']['frac'], d['ctc']['points'][0]['utts_per_s'], d['cpu_baseline'])"; tail -2 gpurun_out/r02_bench_c3_1gpu.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_c3_reference_arm.json 2> /dev/null; tail -c 400 gpurun_out/r02_bench_c3_reference_arm.json
for wl in c1 c2; do python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-ctc > gpurun_out/r02_bench_$wl.json 2> gpurun_out/r02_bench_$wl.err; done
python tools/bench_infer.py > gpurun_out/r02_bench_c5_infer.json 2> gpurun_out/r02_bench_c5_infer.err
python tools/gpu_rec_alone.py > gpurun_out/r02_rec_alone_final.jsonl 2> gpurun_out/r02_rec_alone.err; cat gpurun_out/r02_rec_alone_final.jsonl
python tools/gpu_timeline.py c3 > gpurun_out/r02_timeline_c3.txt 2>&1; head -2 gpurun_out/r02_timeline_c3.txt
python -c "
import __graft_entry__ as g; g.smoke(); print('smoke ok')"
