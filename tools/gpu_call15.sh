#!/bin/bash
# paired 16-utterance sub-groups, forward + BPTT: parity, probes, A/B, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4
for pair in 0 1; do
  echo "== LCB_REC_PAIR=$pair"
  LCB_REC_PAIR=$pair timeout 200 python tools/gpu_rec_profile.py 512 64 1500 > gpurun_out/recprobe_fwd_pair$pair.txt 2>&1; head -2 gpurun_out/recprobe_fwd_pair$pair.txt | tail -1
  LCB_REC_PAIR=$pair timeout 200 python tools/gpu_rec_profile_bwd.py 512 64 1500 > gpurun_out/recprobe_bwd_pair$pair.txt 2>&1; cat gpurun_out/recprobe_bwd_pair$pair.txt
  LCB_REC_PAIR=$pair timeout 200 python tools/gpu_rec_insitu.py 1500 2>&1 | tee gpurun_out/rec_insitu_pair$pair.txt
  LCB_REC_PAIR=$pair timeout 400 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pair$pair.json 2> gpurun_out/bench_pair$pair.err; cut -c1-200 gpurun_out/bench_pair$pair.json
done
