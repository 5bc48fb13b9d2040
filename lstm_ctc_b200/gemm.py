"""16-bit GEMM (tcgen05 + TMA) -- thin host wrapper over lcb_gemm16."""
import contextlib
import threading

import torch

from . import _lib

_tls = threading.local()


@contextlib.contextmanager
def grid_cap(max_ctas):
    """Every gemm() issued inside the block (by this thread) passes max_ctas to the launch: the persistent grid of a GEMM that runs
    on a side stream beside a cluster kernel is capped to the SMs that kernel leaves free.  0 / None: the whole device."""
    old = getattr(_tls, "cap", 0)
    _tls.cap = int(max_ctas or 0)
    try:
        yield
    finally:
        _tls.cap = old


def current_cap():
    return getattr(_tls, "cap", 0)

_DT = {torch.float32: 0, torch.bfloat16: 1, torch.float16: 2}


def _ld(t):
    assert t.dim() == 2 and t.stride(1) == 1
    return t.stride(0)


def gemm(A, B, a_layout=0, b_layout=0, out=None, out_dtype=torch.float32, bias=None, accumulate=False, check=False,
         dropout=None, max_ctas=None, out_bf16=None):
    """C[M,N] (+)= op(A) * op(B) + bias.   A, B: bf16 or fp16 (same type).
    a_layout 0: A is [M,K]; 1: A is [K,M].   b_layout 0: B is [N,K]; 1: B is [K,N].
    dropout = (keep_prob, seed, mask_base): inverted dropout fused into the epilogue; C[r, c] uses element
    mask_base + r*ldc + c of the (seed) mask stream of lcb_dropout16.
    out_bf16: a second destination of the same shape and pitch that receives the (fp16) result as bf16 (lcb_gemm16_twin)."""
    L = _lib.lib()
    assert A.is_cuda and B.is_cuda and _DT.get(A.dtype, 0) and _DT.get(B.dtype, 0), (A.dtype, B.dtype)
    M, K = (A.shape[0], A.shape[1]) if a_layout == 0 else (A.shape[1], A.shape[0])
    N, K2 = (B.shape[0], B.shape[1]) if b_layout == 0 else (B.shape[1], B.shape[0])
    assert K == K2, (A.shape, B.shape, a_layout, b_layout)
    if out is None:
        out = torch.empty(M, N, dtype=out_dtype, device=A.device)
    assert out.shape[0] == M and out.shape[1] == N and out.stride(1) == 1
    cap = current_cap() if max_ctas is None else int(max_ctas)
    if out_bf16 is not None:
        assert out.dtype == torch.float16 and out_bf16.dtype == torch.bfloat16 and out_bf16.shape == out.shape
        assert out_bf16.stride(0) == out.stride(0) and out_bf16.stride(1) == 1 and not accumulate
        keep, seed, base = (float(dropout[0]), int(dropout[1]), int(dropout[2])) if dropout is not None else (1.0, 0, 0)
        st = L.lcb_gemm16_twin(M, N, K, _lib.ptr(A), _ld(A), a_layout, _DT[A.dtype], _lib.ptr(B), _ld(B), b_layout, _DT[B.dtype],
                               _lib.ptr(out), out.stride(0), _DT[out.dtype], _lib.ptr(out_bf16), _lib.ptr(bias), 0,
                               keep, seed, base, cap, _lib.stream_ptr())
        _lib.check(st, "lcb_gemm16_twin")
        return out
    if dropout is not None and dropout[0] < 1.0:
        st = L.lcb_gemm16_dropout(M, N, K, _lib.ptr(A), _ld(A), a_layout, _DT[A.dtype], _lib.ptr(B), _ld(B), b_layout, _DT[B.dtype],
                                  _lib.ptr(out), out.stride(0), _DT[out.dtype], _lib.ptr(bias), 1 if accumulate else 0,
                                  float(dropout[0]), int(dropout[1]), int(dropout[2]), cap, _lib.stream_ptr())
        _lib.check(st, "lcb_gemm16_dropout")
        return out
    args = (M, N, K, _lib.ptr(A), _ld(A), a_layout, _DT[A.dtype], _lib.ptr(B), _ld(B), b_layout, _DT[B.dtype],
            _lib.ptr(out), out.stride(0), _DT[out.dtype], _lib.ptr(bias), 1 if accumulate else 0)
    if check:
        st = L.lcb_gemm16_simt_check(*args, _lib.stream_ptr())
    else:
        st = L.lcb_gemm16(*args, cap, _lib.stream_ptr())
    _lib.check(st, "lcb_gemm16")
    return out
