"""C3 training step time against the share of the idle SMs the side-stream GEMMs may use and the forward launch boundaries
(what a forward recurrence on 96 instead of 64 SMs would leave to the projection GEMMs beside it)."""
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
    model.loss_and_grad(x, lens, y, check_labels=False, seq_len_host=HOST[0])
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


HOST = [lens_h]
from lstm_ctc_b200 import _lib
L = _lib.lib()
combos = [("auto", 1.0), ("auto", 0.75), ("auto", 0.55), ("paired", 1.0), ("paired", 0.75), ("paired", 0.55), ("paired", 0.4), ("auto", 1.0), ("paired", 1.0)]
for layout, sc in combos:
    L.lcb_debug_fwd_layout(1 if layout == "paired" else 0)
    HOST[0] = lens_h
    model.enc.fwd_flow_control = True
    model.enc.flow_fracs = [0.3, 0.55, 0.8]
    model.enc.side_sm_scale = [sc, 1.0]
    print(json.dumps({"fwd_layout": layout, "side_sm_scale_fwd": sc, "ms_per_step": round(timed(), 3), "device_error": L.lcb_device_error(0)}), flush=True)
