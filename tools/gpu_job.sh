mkdir -p gpurun_out
timeout 900 python tools/gpu_side_cap.py > gpurun_out/r02_schedule_ab.jsonl 2> gpurun_out/r02_side_cap.err; cat gpurun_out/r02_schedule_ab.jsonl; tail -3 gpurun_out/r02_side_cap.err
