"""GPU parity: tcgen05 GEMM (lcb_gemm_bf16) vs fp32 torch matmul of the same bf16 operands and vs
the on-device CUDA-core checker.  Tolerance: fp32 accumulation of exact bf16 products -- only the
summation order differs: |err| <= 2e-3 * sqrt(K)-scaled bound, checked as rtol 2e-3 on O(sqrt(K)) values."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _ref(A, B, a_layout, b_layout, bias):
    a = A.float() if a_layout == 0 else A.float().t()
    b = B.float().t() if b_layout == 0 else B.float()
    c = a @ b
    if bias is not None:
        c = c + bias
    return c


@pytest.mark.parametrize("a_layout,b_layout", [(0, 0), (0, 1), (1, 0), (1, 1)])
@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (256, 256, 128), (300, 200, 136), (1000, 1288, 520), (64, 72, 1024)])
def test_gemm_layouts(cuda_dev, a_layout, b_layout, M, N, K):
    from lstm_ctc_b200.gemm import gemm
    torch.manual_seed(M + N + K)
    pad = lambda n: (n + 7) // 8 * 8
    A = torch.randn((M, pad(K)) if a_layout == 0 else (K, pad(M)), device=cuda_dev).bfloat16()
    B = torch.randn((N, pad(K)) if b_layout == 0 else (K, pad(N)), device=cuda_dev).bfloat16()
    Av = A[:, :K] if a_layout == 0 else A[:, :M]
    Bv = B[:, :K] if b_layout == 0 else B[:, :N]
    bias = torch.randn(N, device=cuda_dev)
    C = gemm(Av, Bv, a_layout, b_layout, bias=bias)
    Cc = gemm(Av, Bv, a_layout, b_layout, bias=bias, check=True)
    R = _ref(Av, Bv, a_layout, b_layout, bias)
    torch.cuda.synchronize()
    scale = K ** 0.5
    assert (Cc - R).abs().max() < 2e-3 * scale
    assert (C - R).abs().max() < 2e-3 * scale, ((C - R).abs().max().item(), a_layout, b_layout)


def test_gemm_accumulate_and_bf16_out(cuda_dev):
    from lstm_ctc_b200.gemm import gemm
    torch.manual_seed(0)
    A = torch.randn(512, 256, device=cuda_dev).bfloat16()
    B = torch.randn(384, 256, device=cuda_dev).bfloat16()
    C0 = torch.randn(512, 384, device=cuda_dev)
    C = C0.clone()
    gemm(A, B, out=C, accumulate=True)
    R = C0 + A.float() @ B.float().t()
    assert (C - R).abs().max() < 0.05
    Cb = gemm(A, B, out_dtype=torch.bfloat16)
    assert (Cb.float() - A.float() @ B.float().t()).abs().max() < 0.3


def test_gemm_projection_shape(cuda_dev):
    """The hoisted input projection at C1 scale: [B*T, Din] x [Din, 8H]."""
    from lstm_ctc_b200.gemm import gemm
    torch.manual_seed(1)
    X = torch.randn(16 * 700, 640, device=cuda_dev).bfloat16()
    W = (torch.randn(2560, 640, device=cuda_dev) * 0.05).bfloat16()
    C = gemm(X, W)
    R = X.float() @ W.float().t()
    assert (C - R).abs().max() < 2e-2
    from lstm_ctc_b200 import _lib
    assert _lib.lib().lcb_device_error(0) == 0


@pytest.mark.parametrize("a_layout,b_layout", [(0, 0), (0, 1), (1, 0), (1, 1)])
def test_gemm_fp16_operands(cuda_dev, a_layout, b_layout):
    """Forward GEMMs run fp16 x fp16 (same tensor throughput, 3 more mantissa bits than bf16)."""
    from lstm_ctc_b200.gemm import gemm
    torch.manual_seed(5)
    M, N, K = 384, 264, 520
    A = torch.randn((M, K) if a_layout == 0 else (K, M), device=cuda_dev).half()
    B = torch.randn((N, K) if b_layout == 0 else (K, N), device=cuda_dev).half()
    R = _ref(A, B, a_layout, b_layout, None)
    C = gemm(A, B, a_layout, b_layout)
    Cc = gemm(A, B, a_layout, b_layout, check=True)
    Ch = gemm(A, B, a_layout, b_layout, out_dtype=torch.float16)
    assert (Cc - R).abs().max() < 2e-3 * K ** 0.5
    assert (C - R).abs().max() < 2e-3 * K ** 0.5, (C - R).abs().max().item()
    assert (Ch.float() - R).abs().max() < 0.1


def test_gemm_mixed_formats_rejected(cuda_dev):
    """tcgen05 kind::f16 with A=bf16, B=fp16 is an illegal instruction on B200: the ABI refuses it."""
    from lstm_ctc_b200 import _lib
    from lstm_ctc_b200.gemm import gemm
    A = torch.randn(128, 64, device=cuda_dev).bfloat16()
    B = torch.randn(128, 64, device=cuda_dev).half()
    with pytest.raises(_lib.LcbError):
        gemm(A, B)


@pytest.mark.parametrize("accumulate", [False, True])
def test_gemm_split_k_wgrad_shape(cuda_dev, accumulate):
    """wgrad with few output tiles and a huge reduction (K = frames): the kernel splits K over CTAs and
    accumulates partial tiles with fp32 red.add; also through a strided (column-slice) output view."""
    from lstm_ctc_b200.gemm import gemm
    torch.manual_seed(3)
    K, M, N = 24000, 320, 200
    A = (torch.randn(K, M, device=cuda_dev) * 0.1).bfloat16()
    B = (torch.randn(K, N, device=cuda_dev) * 0.1).bfloat16()
    Cfull = torch.randn(M, N + 56, device=cuda_dev)
    C0 = Cfull.clone()
    out = Cfull[:, 8:8 + N]
    gemm(A, B, 1, 1, out=out, accumulate=accumulate)
    R = A.float().t() @ B.float() + (C0[:, 8:8 + N] if accumulate else 0)
    assert (out - R).abs().max() < 5e-3 * (K ** 0.5) * 0.01 + 1e-2
    assert torch.equal(Cfull[:, :8], C0[:, :8]) and torch.equal(Cfull[:, 8 + N:], C0[:, 8 + N:])   # neighbours untouched


@pytest.mark.parametrize("odt", [torch.float16, torch.bfloat16, torch.float32])
def test_fused_dropout_equals_separate_pass(cuda_dev, odt):
    """lcb_gemm16_dropout: the epilogue mask is the (seed, element index) stream of lcb_dropout16 / lcb_dropout_mask --
    written into a column slice of a wider tensor (mask_base = column offset, row stride = ldc), as the BiLSTM layer
    output uses it -- so it is bit-identical to the GEMM followed by the separate dropout pass."""
    from lstm_ctc_b200 import _lib
    from lstm_ctc_b200.gemm import gemm
    d = torch.device("cuda:0")
    torch.manual_seed(5)
    M, N, K, ld, off = 1000, 192, 136, 448, 64
    A = (torch.randn(M, K, device=d) * 0.3).half()
    B = (torch.randn(N, K, device=d) * 0.3).half()
    keep, seed = 0.8, 0x1234567890ABCDEF
    full = torch.zeros(M, ld, device=d, dtype=odt)
    gemm(A, B, 0, 0, out=full[:, off:off + N], dropout=(keep, seed, off))
    ref = torch.zeros(M, ld, device=d, dtype=odt)
    gemm(A, B, 0, 0, out=ref[:, off:off + N])
    mask = torch.empty(M * ld, dtype=torch.uint8, device=d)
    _lib.check(_lib.lib().lcb_dropout_mask(_lib.ptr(mask), M * ld, keep, seed, _lib.stream_ptr()), "mask")
    mask = mask.view(M, ld).bool()
    want = torch.where(mask, ref.float() / keep, torch.zeros((), device=d)).to(odt)
    torch.cuda.synchronize()
    got, want = full[:, off:off + N].float(), want[:, off:off + N].float()
    assert 0.75 < mask.float().mean().item() < 0.85
    # the fused epilogue scales in fp32 and rounds once; the two-pass reference rounds, scales and rounds again
    tol = {torch.float32: 1e-6, torch.float16: 2e-3, torch.bfloat16: 1.2e-2}[odt]
    assert (got - want).abs().max().item() <= tol * max(1.0, want.abs().max().item())
    assert torch.equal(got == 0, want == 0)
    assert full[:, :off].abs().max().item() == 0 and full[:, off + N:].abs().max().item() == 0


@pytest.mark.parametrize("keep", [1.0, 0.8])
@pytest.mark.parametrize("M,N,K", [(1000, 512, 512), (333, 200, 136), (4096, 64, 64)])
def test_twin_store_writes_the_same_tile_as_bf16(cuda_dev, M, N, K, keep):
    """lcb_gemm16_twin: the fp16 result and its bf16 twin come from the same fp32 accumulator values (bias and dropout applied) --
    each equals the round-to-nearest cast of the fp32 result of the identical launch with fp32 output; a column slice of a wider
    matrix (how the two directions' output projections write their halves) leaves the other columns untouched."""
    from lstm_ctc_b200.gemm import gemm
    torch.manual_seed(M + K)
    A = (torch.randn(M, K, device=cuda_dev) * 0.5).half()
    W = (torch.randn(N, K, device=cuda_dev) * 0.1).half()
    bias = torch.randn(N, device=cuda_dev)
    ld = 2 * ((N + 7) // 8 * 8) + 8
    C16 = torch.full((M, ld), 7.0, device=cuda_dev, dtype=torch.float16)
    Cbf = torch.full((M, ld), 7.0, device=cuda_dev, dtype=torch.bfloat16)
    C32 = torch.zeros((M, ld), device=cuda_dev)
    drop = (keep, 1234, 8) if keep < 1.0 else None
    gemm(A, W, 0, 0, out=C16[:, 8:8 + N], bias=bias, dropout=drop, out_bf16=Cbf[:, 8:8 + N])
    gemm(A, W, 0, 0, out=C32[:, 8:8 + N], bias=bias, dropout=drop)
    torch.cuda.synchronize()
    assert torch.equal(C16[:, 8:8 + N], C32[:, 8:8 + N].half())
    assert torch.equal(Cbf[:, 8:8 + N], C32[:, 8:8 + N].bfloat16())
    for C in (C16, Cbf):
        assert (C[:, :8] == 7.0).all() and (C[:, 8 + N:] == 7.0).all()
    if keep < 1.0:
        z = (C32[:, 8:8 + N] == 0).float().mean().item()
        assert abs(z - (1 - keep)) < 0.02
