set -x
mkdir -p gpurun_out
python tools/gpu_rec_alone.py > gpurun_out/r02_rec_alone_gtma.jsonl 2> gpurun_out/r02_rec_alone.err; cat gpurun_out/r02_rec_alone_gtma.jsonl; tail -3 gpurun_out/r02_rec_alone.err
timeout 900 python -m pytest tests/test_blstm_gpu.py tests/test_full_size_gpu.py tests/test_lstm_uni_gpu.py tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -5
python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r02_bench_c3_gtma.json 2> gpurun_out/r02_bench_c3_gtma.err; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c3_gtma.json')); print(d['ms_per_step'], d['e2e']['ms_per_step'], d['clocks'], {k:(v['ms_total']) for k,v in d['kernels'].items()})"
