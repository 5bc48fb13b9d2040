"""CTC lattice layout A/B (DESIGN K4): one CTA per utterance (alpha + beta as two warp groups) against the two sweeps as the two
CTAs of a cluster, at the in-step shape of the C3 workload and a few small-batch shapes.  Interleaved, medians of several rounds.
Usage: python tools/gpu_ctc_layout_ab.py  -> JSON lines"""
import json
import sys

import torch

sys.path.insert(0, ".")
from lstm_ctc_b200.ctc import ctc_loss_grad  # noqa: E402


def timed(fn, it):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / it


def main():
    d = torch.device("cuda:0")
    shapes = [(64, 1500, 72, 187), (64, 700, 72, 100), (32, 1500, 72, 187), (16, 700, 500, 100), (74, 700, 72, 40), (8, 3000, 72, 375),
              (64, 1500, 72, 400)]
    for B, T, V, L in shapes:
        g = torch.Generator().manual_seed(B + T)
        x = (torch.randn(B, T, V, generator=g) * 3).to(d)
        sl = torch.sort(torch.randint(int(0.8 * T), T + 1, (B,), generator=g)).values.to(torch.int32).to(d)
        lab = torch.randint(0, V - 1, (B, L), generator=g).to(d)
        res = {0: [], 1: []}
        for lay in (0, 1):
            ctc_loss_grad(x, lab, sl, check_labels=False, lattice_layout=lay)
        torch.cuda.synchronize()
        for _ in range(5):
            for lay in (0, 1):
                res[lay].append(timed(lambda: ctc_loss_grad(x, lab, sl, check_labels=False, lattice_layout=lay), 4))
        l0, g0 = ctc_loss_grad(x, lab, sl, check_labels=False, lattice_layout=0)
        l1, g1 = ctc_loss_grad(x, lab, sl, check_labels=False, lattice_layout=1)
        med = {k: sorted(v)[len(v) // 2] for k, v in res.items()}
        print(json.dumps({"B": B, "T": T, "V": V, "L": L, "ms_one_cta": med[0], "ms_two_cta": med[1],
                          "max_abs_grad_diff": float((g0 - g1).abs().max()), "max_rel_loss_diff": float(((l0 - l1).abs() / l0.abs()).max())}),
              flush=True)


if __name__ == "__main__":
    main()
