"""In-kernel phase timing of the forward recurrence (clock64 probes of CTA 0)."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from lstm_ctc_b200 import _lib
from lstm_ctc_b200.blstm import BLSTMEncoder, ModelConfig

H = int(sys.argv[1]) if len(sys.argv) > 1 else 320
B = int(sys.argv[2]) if len(sys.argv) > 2 else 16
T = int(sys.argv[3]) if len(sys.argv) > 3 else 200
dev = torch.device("cuda:0")
cfg = ModelConfig({"input_dim": 120, "num_layers": 1, "num_neurons": H, "num_projects": H, "num_targets": 72, "use_peepholes": True,
                   "dropout_rate": 1.0})
enc = BLSTMEncoder(cfg, dev)
from lstm_ctc_b200.model import random_tf_variables
enc.from_tf_dict(random_tf_variables(cfg, 0))
x = torch.randn(B, T, 120, device=dev)
lens = torch.full((B,), T, dtype=torch.int32, device=dev)
L = _lib.lib()
NS = 64
buf = torch.zeros(NS * 16, dtype=torch.int64, device=dev)
enc.forward(x, lens, training=True)
torch.cuda.synchronize()
SGP = int(sys.argv[4]) if len(sys.argv) > 4 else 0
L.lcb_debug_rec_profile(_lib.ptr(buf), NS | (SGP << 16))
enc.forward(x, lens, training=True)
torch.cuda.synchronize()
L.lcb_debug_rec_profile(None, 0)
p = buf.cpu().numpy().reshape(NS, 16)
names = {0: "ctl:start", 1: "ctl:op_ready", 4: "ctl:turn_acquired", 5: "ctl:blk2_ready", 6: "ctl:blk3_ready", 7: "ctl:mma_issued", 2: "ctl:mma_committed", 12: "cmp:fenced",
         8: "cmp:start", 9: "cmp:g_ready", 10: "cmp:mma_done", 11: "cmp:tmem_ld", 12: "cmp:z_read", 15: "cmp:math_done", 13: "cmp:sent_dsmem", 14: "cmp:global_stores"}
print("H=%d B=%d  cluster: MT=%d NC=%d" % (H, B, enc.rec_mt, enc.rec_nc))
steps = range(20, 60)
base = p[:, 0]
print("step period (ctl start to next ctl start): median %.0f cycles" % np.median(np.diff(p[20:60, 0])))
for k in sorted(names):
    d = [p[s, k] - p[s, 0] for s in steps if p[s, k] > 0]
    if d:
        print("%-20s +%6.0f cycles after ctl:start (median)" % (names[k], np.median(d)))
