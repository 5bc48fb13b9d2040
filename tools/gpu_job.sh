mkdir -p gpurun_out
python tools/gpu_rec_alone.py > gpurun_out/r02_rec_alone_rsdirect.jsonl 2> gpurun_out/r02_rec_alone.err; cat gpurun_out/r02_rec_alone_rsdirect.jsonl; tail -3 gpurun_out/r02_rec_alone.err
