"""C3 training step time against the share of the idle SMs the side-stream GEMMs may use (power / L2 headroom of the recurrence)."""
import sys
import json
import torch
sys.path.insert(0, ".")
import bench
from lstm_ctc_b200.model import AcousticModel

w = bench.WORKLOADS["c3"]
dev = torch.device("cuda:0")
model = AcousticModel(bench.nnet_config(w, 0.9), dev, seed=1234)
x_h, lens_h, y_h = bench.synth_batch(w, 777)
x, lens, y = x_h.to(dev), lens_h.to(dev), y_h.to(dev)


def step():
    model.loss_and_grad(x, lens, y, check_labels=False)
    model.optimizer_step("adam", 4e-4, clip_norm=5.0, l2_decay_weight=1e-5)


def timed(n=12):
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        step()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


combos = [(1.0, 1.0), (0.85, 0.85), (0.7, 0.7), (1.0, 0.8), (0.7, 1.0), (0.55, 0.9), (1.0, 1.0)]
for sf, sb in combos:
    model.enc.side_sm_scale = [sf, sb]
    print(json.dumps({"side_sm_scale_fwd": sf, "side_sm_scale_bwd": sb, "ms_per_step": round(timed(), 3)}), flush=True)
