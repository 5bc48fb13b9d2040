"""Diagnostics run on the GPU box (under gpurun): detailed error maps + quick timings.
Writes human-readable text to stdout; the caller redirects into gpurun_out/."""
import sys
import time
import traceback

import numpy as np
import torch

sys.path.insert(0, ".")
from lstm_ctc_b200 import _lib  # noqa: E402


def ev_time(fn, warm=3, it=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    s = torch.cuda.Event(enable_timing=True); e = torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it):
        fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / it


def gemm_diag():
    from lstm_ctc_b200.gemm import gemm
    d = torch.device("cuda:0")
    print("== GEMM layouts ==")
    for (M, N, K) in [(128, 128, 64), (256, 512, 256), (1000, 1288, 520)]:
        for al in (0, 1):
            for bl in (0, 1):
                torch.manual_seed(1)
                A = torch.randn((M, K) if al == 0 else (K, (M + 7) // 8 * 8), device=d).bfloat16()
                B = torch.randn((N, K) if bl == 0 else (K, (N + 7) // 8 * 8), device=d).bfloat16()
                Av = A if al == 0 else A[:, :M]
                Bv = B if bl == 0 else B[:, :N]
                try:
                    C = gemm(Av, Bv, al, bl)
                    torch.cuda.synchronize()
                    a = Av.float() if al == 0 else Av.float().t()
                    b = Bv.float().t() if bl == 0 else Bv.float()
                    R = a @ b
                    err = (C - R).abs()
                    print("M%d N%d K%d a%d b%d: maxerr %.4g  (ref absmax %.3g) dev_err=%d" % (
                        M, N, K, al, bl, err.max().item(), R.abs().max().item(), _lib.lib().lcb_device_error(1)))
                    if err.max().item() > 0.05 * K ** 0.5:
                        bm = err[: (M // 32) * 32, : (N // 32) * 32].reshape(M // 32, 32, N // 32, 32).amax((1, 3))
                        print("  bad 32x32 blocks (rows=m-block):")
                        for r in range(min(bm.shape[0], 8)):
                            print("   ", " ".join("X" if v > 0.05 * K ** 0.5 else "." for v in bm[r, :16].tolist()))
                except Exception:
                    traceback.print_exc()
    print("== GEMM timing ==")
    for (M, N, K, al, bl) in [(96000, 4096, 1024, 0, 0), (96000, 1024, 4096, 0, 1), (1024, 4096, 96000, 1, 1),
                              (44800, 2560, 640, 0, 0), (8192, 8192, 8192, 0, 0)]:
        A = torch.randn((M, K) if al == 0 else (K, M), device=d).bfloat16()
        B = torch.randn((N, K) if bl == 0 else (K, N), device=d).bfloat16()
        C = torch.empty(M, N, device=d)
        ms = ev_time(lambda: gemm(A, B, al, bl, out=C))
        a = A if al == 0 else A.t()
        b = B.t() if bl == 0 else B
        ms_t = ev_time(lambda: torch.matmul(a, b))
        print("M%d N%d K%d a%d b%d: %.3f ms  %.1f TFLOP/s   (torch bf16 matmul %.3f ms %.1f TFLOP/s) dev_err=%d" % (
            M, N, K, al, bl, ms, 2.0 * M * N * K / ms / 1e9, ms_t, 2.0 * M * N * K / ms_t / 1e9, _lib.lib().lcb_device_error(1)))


def ctc_diag():
    from lstm_ctc_b200.ctc import ctc_loss_grad
    d = torch.device("cuda:0")
    print("== CTC timing (B=256) ==")
    g = torch.Generator().manual_seed(0)
    for (T, L, V) in [(100, 10, 30), (700, 80, 72), (1500, 150, 72), (700, 100, 500), (1500, 100, 5000), (3000, 300, 72), (3000, 300, 5000)]:
        B = 256
        if 8.0 * T * B * V > 40e9:
            continue
        x = (torch.randn(B, T, V, generator=g) * 3).to(d)
        sl = torch.randint(int(0.8 * T), T + 1, (B,), generator=g).to(torch.int32).to(d)
        lab = torch.randint(0, V - 1, (B, L), generator=g).to(d)
        try:
            ms = ev_time(lambda: ctc_loss_grad(x, lab, sl, check_labels=False), warm=2, it=5)
            gb = 8.0 * T * B * V / 1e9
            print("T%d L%d V%d: %.3f ms  %.0f utts/s  %.1f GB/s algorithmic (%.1f%% of 6547.5)" % (
                T, L, V, ms, B / ms * 1e3, gb / ms * 1e3, gb / ms * 1e3 / 6547.5 * 100))
        except Exception:
            traceback.print_exc()
        del x


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0), torch.version.cuda)
    which = sys.argv[1:] or ["gemm", "ctc"]
    for w in which:
        try:
            globals()[w + "_diag"]()
        except Exception:
            traceback.print_exc()
        sys.stdout.flush()
