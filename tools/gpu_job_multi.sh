# N-GPU bench line (gpurun --gpus N -- 'bash tools/gpu_job_multi.sh N')
N=${1:-2}
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/r02_bench_c3_${N}gpu.json 2> gpurun_out/r02_bench_c3_${N}gpu.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c3_${N}gpu.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(json.dumps(d.get('dp_check'))); print(d.get('strong_scaling'))"; tail -3 gpurun_out/r02_bench_c3_${N}gpu.err
