"""Where the data-parallel overhead of a step goes: CUDA events after loss_and_grad (gradients enqueued), after
reducer.finish() (all buckets reduced) and after the optimizer, per rank.  torchrun --nproc-per-node N tools/dp_probe.py"""
import sys
import torch
sys.path.insert(0, ".")
import bench  # noqa: E402
from lstm_ctc_b200 import dist as lcb_dist  # noqa: E402
from lstm_ctc_b200.model import AcousticModel  # noqa: E402

rank, world, device = lcb_dist.init_from_env()
w = bench.WORKLOADS["c3"]
model = AcousticModel(bench.nnet_config(w, 0.9), device, seed=1234)
red = lcb_dist.GradientAllReducer(model.params)
red.broadcast_weights()
x, lens, y = [t.to(device) for t in bench.synth_batch(w, 777 + rank)]
ev = lambda: torch.cuda.Event(enable_timing=True)
acc = [0.0, 0.0, 0.0]
N = 12
for it in range(N + 4):
    e = [ev() for _ in range(4)]
    e[0].record()
    red.begin_step()
    model.loss_and_grad(x, lens, y, bucket_ready=red.bucket_ready, check_labels=False)
    e[1].record()
    red.finish()
    e[2].record()
    model.optimizer_step("adam", 4e-4, clip_norm=5.0, l2_decay_weight=1e-5)
    e[3].record()
    torch.cuda.synchronize()
    if it >= 4:
        for k in range(3):
            acc[k] += e[k].elapsed_time(e[k + 1]) / N
print("rank %d/%d  loss_and_grad %.3f ms  finish(all-reduce tail) %.3f ms  optimizer %.3f ms  total %.3f" % (rank, world, acc[0], acc[1], acc[2], sum(acc)), flush=True)
