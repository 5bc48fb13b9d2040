# final-visit evidence pass (rewritten per gpurun visit; outputs land in gpurun_out/ and are copied to profiles/ by hand)
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-ctc > /dev/null 2>&1; wc -l gpurun_out/r02_launches.csv
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:'lstm_rec|ctc_|gemm_bf16|out_fwd|mos_bwd' -s 1000 -c 360 --csv --log-file gpurun_out/r02_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-ctc > /dev/null 2>&1; wc -l gpurun_out/r02_traffic.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'lstm_rec_fwd2' -s 12 -c 1 -o gpurun_out/r02_full_recfwd -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-ctc > /dev/null 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'lstm_rec_bwd3' -s 40 -c 1 -o gpurun_out/r02_full_recbwd -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-ctc > /dev/null 2>&1
for k in recfwd recbwd; do ncu -i gpurun_out/r02_full_$k.ncu-rep --page raw --csv > gpurun_out/r02_ncu_full_${k}_raw.csv 2>/dev/null; done
ls -la gpurun_out | grep "r02_full\|r02_ncu_full\|r02_traffic\|r02_launches"
