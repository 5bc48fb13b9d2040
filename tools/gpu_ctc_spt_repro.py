"""Repro / check of the coarse CTC lattice mappings (SPT = 8, 16) against the oracle.  argv: spt L T [layout]"""
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
import oracle  # noqa: E402
from lstm_ctc_b200 import _lib  # noqa: E402
from lstm_ctc_b200.ctc import ctc_loss_grad  # noqa: E402

spt, L, T = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
lay = int(sys.argv[4]) if len(sys.argv) > 4 else 0
B, V = 2, 30
_lib.lib().lcb_debug_ctc_min_spt(spt)
rng = np.random.RandomState(0)
x = (rng.randn(B, T, V) * 3).astype(np.float32)
sl = np.array([T, T - 3])
lab = rng.randint(0, V - 1, size=(B, L)).astype(np.int64)
d = torch.device("cuda:0")
loss, grad = ctc_loss_grad(torch.tensor(x, device=d), torch.tensor(lab, device=d), torch.tensor(sl, dtype=torch.int32, device=d), lattice_layout=lay)
torch.cuda.synchronize()
ol, og = oracle.ctc_loss_grad(x.astype(np.float64), lab, sl)
print("spt", spt, "L", L, "T", T, "layout", lay, "loss", loss.cpu().numpy(), ol, "max grad err", np.abs(grad.cpu().numpy() - og).max())
