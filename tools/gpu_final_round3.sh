#!/bin/bash
# Final visit (early-rows backward schedule, vectorised mixture backward, closer CTC plans): parity tests, the bench lines that
# go to profiles/, ncu evidence.  Recurrence kernels are unchanged since gpu_final_round2.sh, so their probes are not repeated.
mkdir -p gpurun_out
S=$(date +%s); t() { echo "[+$(( $(date +%s) - S ))s] $*"; }
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt; tail -3 gpurun_out/pytest_gpu.txt
t pytest
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_c3_1gpu.json 2> gpurun_out/bench_c3_1gpu.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_c3_1gpu.json
t bench
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_c3_reference_arm.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
t ref
for wl in c1 c2; do timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl rc=$?"; done
t c1c2
timeout 600 python tools/bench_infer.py > gpurun_out/bench_c5_infer.json 2> gpurun_out/bench_c5_infer.err; echo "infer rc=$?"
t infer
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
t ncu_list
GM="gpu__time_duration.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,sm__throughput.avg.pct_of_peak_sustained_elapsed"
timeout 900 ncu --metrics $GM --clock-control none -k regex:gemm_bf16_tcgen05 -s 390 -c 200 --csv --log-file gpurun_out/gemm_launch_metrics.csv $B > gpurun_out/ncu_gemm_metrics.log 2>&1; echo "ncu gemm metrics rc=$?"
t ncu_gemm
for spec in "mosbwd:mos_bwd_v4:1:1" "ctc:ctc_alpha_beta:1:1" "gemm:gemm_bf16_tcgen05:200:10"; do
  IFS=: read name rx skip cnt <<< "$spec"
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -o gpurun_out/prof_$name -f $B > gpurun_out/ncu_full_$name.log 2>&1; echo "ncu full $name rc=$?"
  ncu -i gpurun_out/prof_$name.ncu-rep --page raw --csv > gpurun_out/ncu_full_${name}_raw.csv 2>/dev/null
  rm -f gpurun_out/prof_$name.ncu-rep
  t ncu_$name
done
timeout 300 python tools/gpu_timeline.py > gpurun_out/timeline_c3.txt 2>&1
t timeline
timeout 900 python tools/ctc_sweep.py > gpurun_out/ctc_sweep_b256.jsonl 2> gpurun_out/ctc_sweep.err; echo "sweep rc=$?"
t sweep
ls gpurun_out | wc -l
