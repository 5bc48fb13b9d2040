"""Recurrence cost at the full C3 sequence length (HBM regime): event-timed forward (training / inference) and BPTT of ONE layer."""
import sys
import torch
sys.path.insert(0, ".")
from lstm_ctc_b200 import _lib
from lstm_ctc_b200.blstm import BLSTMEncoder, ModelConfig
from lstm_ctc_b200.model import random_tf_variables

H, B, T = 512, 64, int(sys.argv[1]) if len(sys.argv) > 1 else 1500
dev = torch.device("cuda:0")
cfg = ModelConfig({"input_dim": 120, "num_layers": 1, "num_neurons": H, "num_projects": H, "num_targets": 72, "use_peepholes": True, "dropout_rate": 1.0})
enc = BLSTMEncoder(cfg, dev)
enc.from_tf_dict(random_tf_variables(cfg, 0))
x = torch.randn(B, T, 120, device=dev)
lens = torch.full((B,), T, dtype=torch.int32, device=dev)
dX = torch.randn(T * B, 2 * H, device=dev).bfloat16()


def timed(fn, n=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


for hf in (0.3, 0.0):
    enc.head_frac = hf
    print("head_frac %.1f fwd training  %.3f ms (%.0f cycles/step at 1.965 GHz)" % (hf, timed(lambda: enc.forward(x, lens, training=True)), 0))
print("fwd inference %.3f ms" % timed(lambda: enc.forward(x, lens, training=False)))
enc.head_frac = 0.3
enc.forward(x, lens, training=True)
enc.params.gflat.zero_()
print("bwd           %.3f ms" % timed(lambda: enc.backward(dX)))
