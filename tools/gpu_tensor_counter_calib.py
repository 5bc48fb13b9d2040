"""Calibration of ncu's tensor-pipe counters on tcgen05 kernels: one cuBLAS bf16 GEMM (torch.matmul) and one
lcb_gemm16 launch of the SAME 8192^3 problem, to be captured with

  ncu --metrics sm__inst_executed_pipe_tensor.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,\
sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_elapsed,gpu__time_duration.sum,sm__cycles_elapsed.max \
      --clock-control none python tools/gpu_tensor_counter_calib.py

Both kernels deliver about the same TFLOP/s (printed, CUDA events, un-profiled loop first); if the counter read the true
utilisation it would read about the same for both, and about achieved / peak."""
import sys
import torch
sys.path.insert(0, ".")
from lstm_ctc_b200.gemm import gemm  # noqa: E402

d = torch.device("cuda:0")
N = 8192
A = torch.randn(N, N, device=d).bfloat16()
B = torch.randn(N, N, device=d).bfloat16()
C = torch.empty(N, N, device=d, dtype=torch.bfloat16)


def timed(fn, it=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(it):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / it


ms_c = timed(lambda: torch.matmul(A, B.t(), out=C))
ms_o = timed(lambda: gemm(A, B, 0, 0, out=C))
fl = 2.0 * N ** 3
print("cuBLAS (torch.matmul) bf16 8192^3: %.3f ms = %.0f TFLOP/s;  lcb_gemm16: %.3f ms = %.0f TFLOP/s" % (ms_c, fl / ms_c / 1e9, ms_o, fl / ms_o / 1e9))
