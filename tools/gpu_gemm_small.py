"""Small-shape GEMM timing on the GPU box: fixed cost vs per-k-block cost of lcb_gemm16 on the weight-folding shapes."""
import sys
import torch
sys.path.insert(0, ".")
from lstm_ctc_b200.gemm import gemm  # noqa: E402
from tools.gpu_diag import ev_time  # noqa: E402

d = torch.device("cuda:0")
shapes = [  # M, N, K, a_layout, b_layout, out dtype, accumulate
    (2048, 512, 512, 0, 1, torch.float32, False), (2048, 512, 512, 0, 1, torch.float32, True),
    (512, 2048, 512, 1, 0, torch.float32, False), (512, 2048, 512, 1, 0, torch.float32, True),
    (2048, 512, 512, 0, 0, torch.float32, False), (512, 512, 2048, 1, 1, torch.float32, True),
    (512, 512, 2048, 1, 1, torch.float32, False), (512, 512, 8192, 1, 1, torch.float32, False),
    (2048, 512, 2048, 0, 0, torch.float32, False), (2048, 512, 8192, 0, 0, torch.float32, False),
    (96000, 512, 512, 0, 0, torch.float16, False), (96000, 512, 512, 0, 1, torch.float32, False),
    (96000, 4096, 1024, 0, 0, torch.float32, False), (96000, 4096, 1024, 0, 0, torch.float16, False),
    (96000, 4096, 120, 0, 0, torch.float32, False), (96000, 1024, 4096, 0, 1, torch.bfloat16, False),
    (16384, 584, 1024, 0, 0, torch.float32, False), (16384, 1024, 584, 0, 1, torch.bfloat16, False),
]
for (M, N, K, al, bl, odt, acc) in shapes:
    A = torch.randn((M, K) if al == 0 else (K, M), device=d).bfloat16()
    B = torch.randn((N, K) if bl == 0 else (K, N), device=d).bfloat16()
    C = torch.zeros(M, N, device=d, dtype=odt)
    ms = ev_time(lambda: gemm(A, B, al, bl, out=C, accumulate=acc), warm=3, it=20)
    a = A if al == 0 else A.t()
    b = B.t() if bl == 0 else B
    ms_t = ev_time(lambda: torch.matmul(a, b), warm=3, it=20)
    print("M%-6d N%-5d K%-6d a%d b%d out=%-8s acc=%d : %8.1f us  %7.1f TF/s   (torch.matmul bf16 %8.1f us)" % (
        M, N, K, al, bl, str(odt).split('.')[-1], acc, ms * 1e3, 2.0 * M * N * K / ms / 1e9, ms_t * 1e3))
