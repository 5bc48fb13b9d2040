#!/bin/bash
# Final visit of the round (after the uniform-issue / paired sub-group rework): parity tests, the bench lines that go to
# profiles/, ncu evidence for the kernels DESIGN.md quotes.  The CTC sweep (kernel unchanged) runs last.
mkdir -p gpurun_out
S=$(date +%s); t() { echo "[+$(( $(date +%s) - S ))s] $*"; }
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.txt 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.txt; tail -3 gpurun_out/pytest_gpu.txt
t pytest
timeout 900 python bench.py > gpurun_out/bench_c3_1gpu.json 2> gpurun_out/bench_c3_1gpu.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/bench_c3_1gpu.json
t bench
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_c3_reference_arm.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
t ref
for wl in c1 c2; do timeout 600 python bench.py --workload $wl --steps 10 --warmup 3 > gpurun_out/bench_$wl.json 2> gpurun_out/bench_$wl.err; echo "bench $wl rc=$?"; done
t c1c2
timeout 600 python tools/bench_infer.py > gpurun_out/bench_c5_infer.json 2> gpurun_out/bench_c5_infer.err; echo "infer rc=$?"
t infer
B="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/launches.csv $B > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu list rc=$?"
t ncu_list
GM="gpu__time_duration.sum,sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum,dram__bytes_read.sum,dram__bytes_write.sum,launch__grid_size,sm__throughput.avg.pct_of_peak_sustained_elapsed"
timeout 900 ncu --metrics $GM --clock-control none -k regex:gemm_bf16_tcgen05 -s 240 -c 130 --csv --log-file gpurun_out/gemm_launch_metrics.csv $B > gpurun_out/ncu_gemm_metrics.log 2>&1; echo "ncu gemm metrics rc=$?"
t ncu_gemm
for spec in "gemm:gemm_bf16_tcgen05:150:10" "recfwd:lstm_rec_fwd2:6:2" "recbwd:lstm_rec_bwd3:3:1" "ctc:ctc_alpha_beta:1:1" "ctcsm:ctc_softmax:1:1" "outfwd:out_fwd_kernel:1:1" "mosbwd:mos_bwd_kernel:1:1" "optim:apply_update:1:1"; do
  IFS=: read name rx skip cnt <<< "$spec"
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$rx -s $skip -c $cnt -o gpurun_out/prof_$name -f $B > gpurun_out/ncu_full_$name.log 2>&1; echo "ncu full $name rc=$?"
  ncu -i gpurun_out/prof_$name.ncu-rep --page raw --csv > gpurun_out/ncu_full_${name}_raw.csv 2>/dev/null
  rm -f gpurun_out/prof_$name.ncu-rep
  t ncu_$name
done
python tools/ncu_summary.py gemm=gpurun_out/ncu_full_gemm_raw.csv recfwd=gpurun_out/ncu_full_recfwd_raw.csv recbwd=gpurun_out/ncu_full_recbwd_raw.csv ctc=gpurun_out/ncu_full_ctc_raw.csv ctcsm=gpurun_out/ncu_full_ctcsm_raw.csv outfwd=gpurun_out/ncu_full_outfwd_raw.csv mosbwd=gpurun_out/ncu_full_mosbwd_raw.csv optim=gpurun_out/ncu_full_optim_raw.csv > gpurun_out/ncu_full_summary.txt 2>&1
for sgp in 0 1; do timeout 200 python tools/gpu_rec_profile.py 512 64 1500 $sgp > gpurun_out/recprobe_fwd_sg$sgp.txt 2>&1; timeout 200 python tools/gpu_rec_profile_bwd.py 512 64 1500 $sgp > gpurun_out/recprobe_bwd_sg$sgp.txt 2>&1; done
LCB_REC_PAIR=0 timeout 200 python tools/gpu_rec_profile.py 512 64 1500 > gpurun_out/recprobe_fwd_bg32.txt 2>&1
timeout 200 python tools/gpu_rec_profile.py 512 32 1500 > gpurun_out/recprobe_fwd_bg16_alone.txt 2>&1
timeout 300 python tools/gpu_timeline.py > gpurun_out/timeline_c3.txt 2>&1
timeout 300 python tools/gpu_gemm_small.py > gpurun_out/gemm_shapes.txt 2>&1
t probes
timeout 900 python tools/ctc_sweep.py > gpurun_out/ctc_sweep_b256.jsonl 2> gpurun_out/ctc_sweep.err; echo "sweep rc=$?"
t sweep
ls gpurun_out | wc -l
