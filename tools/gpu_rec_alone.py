"""Recurrence kernels alone (no GEMM beside them), event-timed at the C3 cell size: time per scan step against the batch --
16-utterance groups alone in their clusters (B <= 48: 2 x 3 clusters resident) or paired (B = 64), full T."""
import sys
import json
import torch
sys.path.insert(0, ".")
from lstm_ctc_b200 import _lib

L = _lib.lib()
dev = torch.device("cuda:0")
Hp = 512
T = int(sys.argv[1]) if len(sys.argv) > 1 else 1500
g = torch.Generator(device="cpu").manual_seed(0)


def run(B, lens_mode, host_lens=False):
    N = T * B
    G = (torch.randn(N, 8 * Hp, generator=g) * 0.5).to(dev)
    fold16 = (torch.randn(8 * Hp, Hp, generator=g) * 0.03).to(dev).half()
    foldb = (torch.randn(2 * Hp, 4 * Hp, generator=g) * 0.03).to(dev).bfloat16()
    peep = (torch.randn(2, 3, Hp, generator=g) * 0.1).to(dev)
    if lens_mode == "full":
        lens = torch.full((B,), T, dtype=torch.int32)
    else:
        lens = torch.sort(torch.randint(int(0.8 * T), T + 1, (B,), generator=g).int()).values
        lens[-1] = T
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
    st = _lib.stream_ptr()

    lens_host = (_lib.ctypes.c_int32 * B)(*[int(v) for v in lens.cpu().tolist()]) if host_lens else None

    def fwd():
        _lib.check(L.lcb_lstm_rec_fwd_range_hl(_lib.ptr(G), _lib.ptr(fold16), _lib.ptr(peep), _lib.ptr(lens), lens_host, None, _lib.ptr(M),
                                               _lib.ptr(gates), _lib.ptr(cst), None, None, T, B, Hp, 2, 5.0, 0, T, _lib.ptr(ws), ws.numel(),
                                               st), "fwd")

    def bwd():
        _lib.check(L.lcb_lstm_rec_bwd_range(_lib.ptr(dM), _lib.ptr(gates), _lib.ptr(cst), _lib.ptr(foldb), _lib.ptr(peep), _lib.ptr(lens),
                                            _lib.ptr(dG), _lib.ptr(dbias), _lib.ptr(dpeep), T, B, Hp, 2, 0, T, _lib.ptr(carry),
                                            _lib.ptr(ws), ws.numel(), st), "bwd")

    out = {"B": B, "T": T, "lens": lens_mode, "host_lens": host_lens, "grid_fwd": L.lcb_lstm_rec_grid(B, Hp, 2, 0), "grid_bwd": L.lcb_lstm_rec_grid(B, Hp, 2, 1)}
    for name, fn in (("fwd", fwd), ("bwd", bwd)):
        fn(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        out[name + "_ms"] = round(ms, 3)
        out[name + "_us_per_step"] = round(ms * 1e3 / T, 3)
    out["device_error"] = L.lcb_device_error(0)
    print(json.dumps(out), flush=True)


for B in (16, 48, 64):
    run(B, "full")
run(64, "sorted_0.8T..T")
run(64, "sorted_0.8T..T", host_lens=True)
if hasattr(L, "lcb_debug_fwd_layout"):
    L.lcb_debug_fwd_layout(1)
    print("# forward layout forced to two groups per cluster (round-1 layout)")
    run(64, "full")
    run(64, "sorted_0.8T..T", host_lens=True)
    L.lcb_debug_fwd_layout(0)
