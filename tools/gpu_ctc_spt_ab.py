"""CTC lattice: states per thread (SPT) and layout against batch size / label length -- does a coarser thread mapping that keeps every
utterance's CTA resident in ONE wave beat the finest mapping in two waves?  JSON lines: ms per call by (spt, layout)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from lstm_ctc_b200 import _lib  # noqa: E402
from lstm_ctc_b200.ctc import ctc_loss_grad  # noqa: E402


def timed(fn, it=4):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / it


d = torch.device("cuda:0")
L_ = _lib.lib()
for B, T, V, L in [(256, 700, 72, 100), (256, 700, 72, 300), (256, 700, 500, 100), (256, 700, 500, 300), (256, 700, 5000, 300), (256, 700, 5000, 100),
                   (64, 1500, 72, 187), (128, 700, 72, 100), (256, 1500, 72, 187)]:
    g = torch.Generator().manual_seed(B + T)
    x = (torch.randn(B, T, V, generator=g) * 3).to(d)
    sl = torch.randint(int(0.8 * T), T + 1, (B,), generator=g).to(torch.int32).to(d)
    lab = torch.randint(0, V - 1, (B, L), generator=g).to(d)
    rec = {"B": B, "T": T, "V": V, "L": L}
    for spt in (2, 4, 8):
        L_.lcb_debug_ctc_min_spt(spt)
        for lay in (0, 1):
            fn = lambda: ctc_loss_grad(x, lab, sl, check_labels=False, lattice_layout=lay)
            print("# running", B, T, V, L, spt, lay, file=sys.stderr, flush=True)
            fn(); torch.cuda.synchronize()
            rec["spt%d_%s" % (spt, "two_cta" if lay else "one_cta")] = round(min(timed(fn) for _ in range(3)), 4)
    L_.lcb_debug_ctc_min_spt(2)
    print(json.dumps(rec), flush=True)
