"""Power draw of the step's components run alone, ~2.5 s each, sampled with nvidia-smi (100 ms): the forward recurrence, BPTT, a bulk GEMM
on the whole chip and on the SMs the recurrence leaves free, the CTC lattice, recurrence + GEMM together.  The C3 step runs at the
1000 W cap (tools/gpu_power_probe.sh); this shows which component the energy goes to.  JSON lines."""
import json
import subprocess
import sys
import threading
import time

import torch

sys.path.insert(0, ".")
from lstm_ctc_b200 import _lib
from lstm_ctc_b200.gemm import gemm
from lstm_ctc_b200.ctc import ctc_loss_grad

L = _lib.lib()
dev = torch.device("cuda:0")
Hp, T, B = 512, 1500, 64
g = torch.Generator(device="cpu").manual_seed(0)
N = T * B
G = (torch.randn(N, 8 * Hp, generator=g) * 0.5).to(dev)
fold16 = (torch.randn(8 * Hp, Hp, generator=g) * 0.03).to(dev).half()
foldb = (torch.randn(2 * Hp, 4 * Hp, generator=g) * 0.03).to(dev).bfloat16()
peep = (torch.randn(2, 3, Hp, generator=g) * 0.1).to(dev)
lens = torch.sort(torch.randint(int(0.8 * T), T + 1, (B,), generator=g).int()).values
lens[-1] = T
lens_host = (_lib.ctypes.c_int32 * B)(*[int(v) for v in lens.tolist()])
lens = lens.to(dev)
M = torch.empty(N, 2 * Hp, dtype=torch.float16, device=dev)
gates = torch.empty(N, 2 * Hp, dtype=torch.int64, device=dev)
cst = torch.empty(N, 2 * Hp, dtype=torch.float32, device=dev)
ws = torch.empty(max(16, L.lcb_lstm_rec_workspace_bytes(B, Hp)), dtype=torch.uint8, device=dev)
dM = (torch.randn(N, 2 * Hp, generator=g) * 0.01).to(dev)
dG = torch.empty(N, 8 * Hp, dtype=torch.bfloat16, device=dev)
dbias = torch.zeros(8 * Hp, device=dev)
dpeep = torch.zeros(2, 3, Hp, device=dev)
carry = torch.zeros(B * 2 * Hp * 2, device=dev)
X16 = (torch.randn(N, 1024, generator=g)).to(dev).half()
W16 = (torch.randn(4096, 1024, generator=g) * 0.03).to(dev).half()
Gout = torch.empty(N, 4096, dtype=torch.float32, device=dev)
xl = (torch.randn(B, T, 72, generator=g) * 3).to(dev)
lab = torch.randint(0, 71, (B, 187), generator=g).to(dev)
side = torch.cuda.Stream()


def fwd():
    _lib.check(L.lcb_lstm_rec_fwd_range_hl(_lib.ptr(G), _lib.ptr(fold16), _lib.ptr(peep), _lib.ptr(lens), lens_host, None, _lib.ptr(M),
                                           _lib.ptr(gates), _lib.ptr(cst), None, None, T, B, Hp, 2, 5.0, 0, T, _lib.ptr(ws), ws.numel(),
                                           _lib.stream_ptr()), "fwd")


def bwd():
    _lib.check(L.lcb_lstm_rec_bwd_range(_lib.ptr(dM), _lib.ptr(gates), _lib.ptr(cst), _lib.ptr(foldb), _lib.ptr(peep), _lib.ptr(lens),
                                        _lib.ptr(dG), _lib.ptr(dbias), _lib.ptr(dpeep), T, B, Hp, 2, 0, T, _lib.ptr(carry),
                                        _lib.ptr(ws), ws.numel(), _lib.stream_ptr()), "bwd")


def gemm_cap(cap):
    return lambda: gemm(X16, W16, 0, 0, out=Gout, max_ctas=cap)


def both(rec, cap):
    def f():
        rec()
        with torch.cuda.stream(side):
            for _ in range(6):
                gemm(X16, W16, 0, 0, out=Gout, max_ctas=cap)
    return f


samples = []
stop = False


def sampler():
    p = subprocess.Popen(["nvidia-smi", "--query-gpu=power.draw.instant,clocks.sm", "--format=csv,noheader,nounits", "-lms", "100"],
                         stdout=subprocess.PIPE, text=True)
    while not stop:
        line = p.stdout.readline()
        if not line:
            break
        try:
            w, c = line.strip().split(", ")
            samples.append((time.time(), float(w), float(c)))
        except ValueError:
            pass
    p.kill()


th = threading.Thread(target=sampler, daemon=True)
th.start()
time.sleep(1.5)
fwd(); bwd(); torch.cuda.synchronize()
phases = [("idle", None), ("forward recurrence alone (96 SMs)", fwd), ("BPTT alone (64 SMs)", bwd), ("GEMM 96000x4096x1024 fp32 out, 148 SMs", gemm_cap(0)),
          ("GEMM, 52 SMs", gemm_cap(52)), ("GEMM, 84 SMs", gemm_cap(84)), ("CTC loss+grad in-step shape", lambda: ctc_loss_grad(xl, lab, lens, check_labels=False)),
          ("forward recurrence + GEMMs on 52 SMs", both(fwd, 52)), ("BPTT + GEMMs on 84 SMs", both(bwd, 84))]
for name, fn in phases:
    t0 = time.time()
    n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.time() - t0 < 2.5:
        if fn is None:
            time.sleep(0.1)
        else:
            fn()
            n += 1
            if n % 4 == 0:
                torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    t1 = time.time()
    time.sleep(0.3)
    ss = [(w, c) for (ts, w, c) in samples if t0 + 0.6 < ts < t1 - 0.1]
    ws_ = sorted(w for w, _ in ss); cs = sorted(c for _, c in ss)
    print(json.dumps({"phase": name, "calls": n, "ms_per_call": round(e0.elapsed_time(e1) / max(n, 1), 3), "samples": len(ss),
                      "power_w_median": ws_[len(ws_) // 2] if ws_ else None, "power_w_max": ws_[-1] if ws_ else None,
                      "sm_mhz_median": cs[len(cs) // 2] if cs else None}), flush=True)
stop = True
