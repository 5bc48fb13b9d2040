#!/bin/bash
# Round profile visit: ncu launch list of the bench command, full captures of the dominant kernels, C1/C2 bench lines,
# C5 inference bench, C4 CTC sweep.  Everything lands in gpurun_out/ ; summaries are copied to profiles/ by hand.
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_tensor.sum,sm__throughput.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct,sm__warps_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum,sm__pipe_tensor_op_hmma_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_active.avg"
for spec in "gemm:regex:gemm_bf16_tcgen05_kernel<256, 0, 0, 0>:12:4" "recfwd:regex:lstm_rec_fwd2:6:2" "recbwd:regex:lstm_rec_bwd3:3:2" "ctc:regex:ctc_alpha_beta:1:1" "ctcsm:regex:ctc_softmax:1:1" "outfwd:regex:out_fwd_kernel:1:1"; do
  IFS=: read name r1 r2 skip cnt <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k "$r1:$r2" -s $skip -c $cnt -o gpurun_out/prof_$name -f $B > gpurun_out/ncu_full_$name.log 2>&1; echo "ncu full $name rc=$?"
  ncu -i gpurun_out/prof_$name.ncu-rep --page raw --csv > gpurun_out/ncu_full_${name}_raw.csv 2>/dev/null
  ncu -i gpurun_out/prof_$name.ncu-rep --page source --csv 2>/dev/null | gzip > gpurun_out/ncu_full_${name}_source.csv.gz
  if [ "$name" != "gemm" ] && [ "$name" != "recbwd" ]; then rm -f gpurun_out/prof_$name.ncu-rep; fi
done
for wl in c1 c2; do timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl rc=$?"; done
timeout 600 python tools/bench_infer.py > gpurun_out/bench_c5_infer.json 2> gpurun_out/bench_c5_infer.err; echo "infer rc=$?"; cat gpurun_out/bench_c5_infer.json
timeout 1500 python tools/ctc_sweep.py > gpurun_out/ctc_sweep_b256.jsonl 2> gpurun_out/ctc_sweep.err; echo "sweep rc=$?"; tail -3 gpurun_out/ctc_sweep_b256.jsonl
rm -f gpurun_out/prof_*.ncu-rep.tmp
ls -la gpurun_out | head -50
