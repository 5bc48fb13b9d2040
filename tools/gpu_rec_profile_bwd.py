"""In-kernel phase timing of the BPTT kernel (clock64 probes of CTA 0, sub-group 0)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from lstm_ctc_b200 import _lib
from lstm_ctc_b200.blstm import BLSTMEncoder, ModelConfig
from lstm_ctc_b200.model import random_tf_variables

H = int(sys.argv[1]) if len(sys.argv) > 1 else 512
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
T = int(sys.argv[3]) if len(sys.argv) > 3 else 200
dev = torch.device("cuda:0")
cfg = ModelConfig({"input_dim": 120, "num_layers": 1, "num_neurons": H, "num_projects": H, "num_targets": 72, "use_peepholes": True, "dropout_rate": 1.0})
enc = BLSTMEncoder(cfg, dev)
enc.from_tf_dict(random_tf_variables(cfg, 0))
x = torch.randn(B, T, 120, device=dev)
lens = torch.full((B,), T, dtype=torch.int32, device=dev)
L = _lib.lib()
NS = 64
buf = torch.zeros(NS * 16, dtype=torch.int64, device=dev)
out = enc.forward(x, lens, training=True)
dX = torch.randn(T * B, 2 * H, device=dev).bfloat16()
enc.params.gflat.zero_()
enc.backward(dX)
torch.cuda.synchronize()
enc.forward(x, lens, training=True)
torch.cuda.synchronize()
SGP = int(sys.argv[4]) if len(sys.argv) > 4 else 0
L.lcb_debug_rec_profile(_lib.ptr(buf), NS | (SGP << 16))
enc.backward(dX)
torch.cuda.synchronize()
L.lcb_debug_rec_profile(None, 0)
p = buf.cpu().numpy().reshape(NS, 16)
names = {0: "iss:start", 1: "iss:dz_ready", 4: "iss:turn_acquired", 7: "iss:mma_issued", 2: "iss:committed", 8: "cmp:start", 9: "cmp:prefetch_issued",
         10: "cmp:red_ready", 11: "cmp:dz_done", 12: "cmp:arrived", 13: "cmp:mma_done", 14: "cmp:sent"}
print("BPTT H=%d B=%d NC=%d" % (H, B, enc.rec_nc))
print("step period: median %.0f cycles" % np.median(np.diff(p[20:60, 8])))
for k in sorted(names):
    d = [p[s, k] - p[s, 8] for s in range(20, 60) if p[s, k] > 0]
    if d:
        print("%-22s +%6.0f cycles after cmp:start (median)" % (names[k], np.median(d)))
