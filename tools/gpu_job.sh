# Evidence pass of a GPU visit (run as: gpurun --timeout 3000 -- 'bash tools/gpu_job.sh'; 2 / 4 / 8-GPU lines: tools/gpu_job_multi.sh under
# gpurun --gpus N).  Outputs land in gpurun_out/; what is quoted in DESIGN.md is copied to profiles/ (index: profiles/README.md).
# The ncu passes at the end serialise kernels: the bench then runs the range-launch schedule (blstm._kernels_serialised), and a number
# printed under ncu is never a bench value.
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/r02_pytest_gpu.txt; cat gpurun_out/r02_pytest_gpu.txt
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_c3_1gpu.json 2> gpurun_out/r02_bench_c3_1gpu.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c3_1gpu.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['clocks'], {k:(v['ms_total']) for k,v in d['kernels'].items()}); print(d['roofline']['us_per_time_step'], d['roofline']['frac'], d['rooflines_other']['gemm']['frac'], d['ctc']['points'][0]['utts_per_s'], d['cpu_baseline'])"; tail -2 gpurun_out/r02_bench_c3_1gpu.err
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_c3_reference_arm.json 2> /dev/null; tail -c 400 gpurun_out/r02_bench_c3_reference_arm.json
for wl in c1 c2; do python bench.py --workload $wl --steps 20 --warmup 5 --no-cpu-baseline --no-ctc > gpurun_out/r02_bench_$wl.json 2> gpurun_out/r02_bench_$wl.err; done
python tools/bench_infer.py > gpurun_out/r02_bench_c5_infer.json 2> gpurun_out/r02_bench_c5_infer.err
python tools/gpu_rec_alone.py > gpurun_out/r02_rec_alone_final.jsonl 2> gpurun_out/r02_rec_alone.err; cat gpurun_out/r02_rec_alone_final.jsonl
python tools/gpu_timeline.py c3 > gpurun_out/r02_timeline_c3.txt 2>&1; head -2 gpurun_out/r02_timeline_c3.txt
timeout 400 python tools/ctc_sweep.py > gpurun_out/r02_ctc_sweep_b256.jsonl 2> gpurun_out/r02_ctc_sweep.err; wc -l gpurun_out/r02_ctc_sweep_b256.jsonl
timeout 120 python tools/gpu_ctc_layout_ab.py > gpurun_out/r02_ctc_layout_ab.jsonl 2>&1
python -c "
import __graft_entry__ as g; g.smoke(); print('smoke ok')"
# ---- ncu: launch census, DRAM traffic per launch, --set full of the CTC lattice (two-CTA layout) and the forward recurrence ----
BENCH="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-ctc"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches.csv $BENCH > /dev/null 2>&1; wc -l gpurun_out/r02_launches.csv
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'lstm_rec|ctc_|gemm_bf16|out_fwd|mos_bwd' -s 1000 -c 360 --csv --log-file gpurun_out/r02_traffic.csv $BENCH > /dev/null 2>&1; wc -l gpurun_out/r02_traffic.csv
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'ctc_lattice' -s 3 -c 1 -o gpurun_out/r02_full_ctclat -f $BENCH > /dev/null 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'lstm_rec_fwd2' -s 12 -c 1 -o gpurun_out/r02_full_recfwd -f $BENCH > /dev/null 2>&1
for k in ctclat recfwd; do ncu -i gpurun_out/r02_full_$k.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_${k}_raw.csv 2>/dev/null; done
ls -la gpurun_out | grep "r02_full\|r02_ncu_full\|r02_traffic\|r02_launches"
