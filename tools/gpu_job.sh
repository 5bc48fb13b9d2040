set -x
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/r02_bench_c3_1gpu.json 2> gpurun_out/r02_bench_c3_1gpu.err; tail -c 600 gpurun_out/r02_bench_c3_1gpu.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-ctc > /dev/null 2>&1
ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'lstm_rec|ctc_|gemm_bf16|out_fwd|mos_bwd' -s 900 -c 330 --csv --log-file gpurun_out/r02_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-ctc > /dev/null 2>&1
ncu --metrics sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,sm__cycles_elapsed.max --clock-control none --csv --log-file gpurun_out/r02_tensor_counter_calib.csv python tools/gpu_tensor_counter_calib.py > gpurun_out/r02_tensor_counter_calib.txt 2>&1
tail -2 gpurun_out/r02_tensor_counter_calib.txt
python tools/ctc_sweep.py > gpurun_out/r02_ctc_sweep_b256.jsonl 2>/dev/null; wc -l gpurun_out/r02_ctc_sweep_b256.jsonl
ls -la gpurun_out
