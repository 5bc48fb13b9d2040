"""C4: CTC loss+grad microbenchmark sweep (B=256): utts/s and achieved algorithmic HBM GB/s (8*T*B*V bytes)
against the measured peak, with the CPU oracle (C port of the TF CPU kernel, all host threads) beside it on a
bounded subset.  Usage: python tools/ctc_sweep.py [--quick] [--one T L V]"""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from lstm_ctc_b200.ctc import ctc_loss_grad  # noqa: E402

PEAK = json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"] if os.path.exists("MEASURED_PEAKS.json") else 6650.0


def run(T, L, V, B=256, iters=5, cpu=False):
    d = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    x = (torch.randn(B, T, V, generator=g) * 3).to(d)
    sl = torch.randint(int(0.8 * T), T + 1, (B,), generator=g).to(torch.int32).to(d)
    lab = torch.randint(0, V - 1, (B, L), generator=g).to(d)
    for _ in range(2):
        ctc_loss_grad(x, lab, sl, check_labels=False)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        ctc_loss_grad(x, lab, sl, check_labels=False)
    e.record(); torch.cuda.synchronize()
    ms = s.elapsed_time(e) / iters
    gb = 8.0 * T * B * V / 1e9
    rec = {"T": T, "L": L, "V": V, "B": B, "ms": ms, "utts_per_s": B / ms * 1e3, "gbs": gb / ms * 1e3,
           "frac_of_hbm_peak": gb / ms * 1e3 / PEAK}
    if cpu:
        import oracle
        xs, ls, ss = x[:32].cpu().numpy(), lab[:32].cpu().numpy(), sl[:32].cpu().numpy()
        t0 = time.perf_counter()
        oracle.ctc_loss_grad(xs, ls, ss, dtype=np.float32, nthreads=os.cpu_count())
        dt = time.perf_counter() - t0
        rec["cpu_utts_per_s"] = 32 / dt
        rec["cpu_cores"] = os.cpu_count()
    return rec


if __name__ == "__main__":
    if "--one" in sys.argv:
        i = sys.argv.index("--one")
        print(json.dumps(run(int(sys.argv[i + 1]), int(sys.argv[i + 2]), int(sys.argv[i + 3]), iters=2)))
        sys.exit(0)
    quick = "--quick" in sys.argv
    Ts = [100, 700, 3000] if quick else [100, 300, 700, 1500, 3000]
    Ls = [10, 100] if quick else [10, 30, 100, 300]
    Vs = [30, 72, 5000] if quick else [30, 72, 500, 5000]
    for T in Ts:
        for L in Ls:
            if L > T // 2:
                continue
            for V in Vs:
                if 8.0 * T * 256 * V > 36e9:
                    continue
                print(json.dumps(run(T, L, V, cpu=(V == 72 and L in (10, 100) and T <= 700))), flush=True)
