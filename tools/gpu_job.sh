mkdir -p gpurun_out
python tools/gpu_rec_alone.py > gpurun_out/r02_rec_alone_ptrs.jsonl 2> gpurun_out/r02_rec_alone.err; cat gpurun_out/r02_rec_alone_ptrs.jsonl; tail -3 gpurun_out/r02_rec_alone.err
timeout 900 python -m pytest tests/test_blstm_gpu.py tests/test_full_size_gpu.py tests/test_lstm_uni_gpu.py tests/test_model_gpu.py -x -q -m gpu 2>&1 | tail -5
