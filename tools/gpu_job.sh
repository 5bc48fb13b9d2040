mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 20 --warmup 5 > gpurun_out/r02_bench_c3_4gpu.json 2> gpurun_out/r02_bench_c3_4gpu.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_c3_4gpu.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print(json.dumps(d.get('dp_check'))); print(d.get('strong_scaling'))"; tail -3 gpurun_out/r02_bench_c3_4gpu.err
