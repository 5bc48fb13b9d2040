"""Probe (one subprocess per case) which 16-bit operand format combinations tcgen05.mma kind::f16 accepts."""
import subprocess
import sys

CASE = '''
import sys, torch
sys.path.insert(0, ".")
from lstm_ctc_b200.gemm import gemm
adt, bdt = {dts}
torch.manual_seed(0)
A = torch.randn(256, 128, device="cuda").to(adt); B = torch.randn(128, 128, device="cuda").to(bdt)
C = gemm(A, B)
torch.cuda.synchronize()
print("OK maxerr", (C - A.float() @ B.float().t()).abs().max().item())
'''
for dts in ("torch.bfloat16, torch.bfloat16", "torch.float16, torch.float16", "torch.bfloat16, torch.float16", "torch.float16, torch.bfloat16"):
    r = subprocess.run([sys.executable, "-c", CASE.format(dts=dts)], capture_output=True, text=True, env={**__import__("os").environ, "CUDA_LAUNCH_BLOCKING": "1"})
    print(dts, "->", (r.stdout.strip().splitlines() or ["(no stdout)"])[-1], "|", (r.stderr.strip().splitlines() or [""])[-1][:160])
