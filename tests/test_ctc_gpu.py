"""GPU parity: lcb_ctc_loss_grad_f32 (through the C ABI) vs the CPU oracle.
Tolerance (north_star): 1e-4 relative on loss and gradient, fp32."""
import json
import os

import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(__file__), "golden")
RTOL = 1e-4


_LAYOUT = [0]


@pytest.fixture(autouse=True, params=[0, 1], ids=["one_cta_per_utt", "two_cta_per_utt"])
def lattice_layout(request):
    """Every oracle comparison runs on both lattice layouts (lcb_ctc_loss_grad_f32_layout): alpha + beta sweeps as two warp
    groups of one CTA, and as the two CTAs of a cluster (what the default picks when 2 B <= SMs)."""
    _LAYOUT[0] = request.param
    yield request.param


def run_gpu(logits, labels, seq_len):
    from lstm_ctc_b200.ctc import ctc_loss_grad
    d = torch.device("cuda:0")
    loss, grad = ctc_loss_grad(torch.tensor(logits, dtype=torch.float32, device=d),
                               torch.tensor(labels, dtype=torch.int64, device=d),
                               torch.tensor(seq_len, dtype=torch.int32, device=d), lattice_layout=_LAYOUT[0])
    torch.cuda.synchronize()
    return loss.cpu().numpy().astype(np.float64), grad.cpu().numpy().astype(np.float64)


def assert_close(loss, grad, oloss, ograd, what):
    fin = np.isfinite(oloss)
    assert np.array_equal(np.isfinite(loss), fin), what
    assert np.allclose(loss[fin], oloss[fin], rtol=RTOL, atol=1e-5), (what, loss, oloss)
    # gradient entries are O(1) probabilities: 1e-4 relative to the per-frame scale (max |g| <= 1)
    err = np.abs(grad - ograd).max()
    assert err < RTOL, (what, err)
    rel = np.abs(grad - ograd) / np.maximum(np.abs(ograd), 1e-2)
    assert rel.max() < 5 * RTOL, (what, rel.max())


def test_tf_known_answers(cuda_dev):
    kat = json.load(open(os.path.join(G, "ctc_tf_kat.json")))
    loss, grad = run_gpu(np.log(np.array(kat["probs"])), kat["labels"], kat["seq_len"])
    assert np.allclose(loss, kat["loss"], atol=5e-5)
    assert np.allclose(grad, np.array(kat["grad"]), atol=5e-6)


def test_golden_cases(cuda_dev):
    for c in json.load(open(os.path.join(G, "ctc_cases.json"))):
        loss, grad = run_gpu(c["logits"], c["labels"], c["seq_len"])
        oloss = np.array([np.inf if v == "inf" else v for v in c["loss"]])
        assert_close(loss, grad, oloss, np.array(c["grad"]), c["desc"])


def _random_case(rng, B, T, V, Lmax, scale=3.0, full_len=False):
    x = (rng.randn(B, T, V) * scale).astype(np.float32)
    sl = np.full(B, T) if full_len else rng.randint(max(1, int(0.6 * T)), T + 1, size=B)
    lab = -np.ones((B, max(Lmax, 1)), dtype=np.int64)
    for b in range(B):
        n = rng.randint(0, Lmax + 1)
        n = min(n, int(sl[b]) // 2)
        lab[b, :n] = rng.randint(0, V - 1, size=n)
    return x, lab, sl


@pytest.mark.parametrize("B,T,V,Lmax", [
    (8, 50, 30, 10),      # 1 warp / 2 states per thread
    (5, 120, 72, 60),     # 2 warps x 2
    (3, 400, 72, 180),    # 6 warps x 2
    (3, 500, 40, 230),    # 8 warps x 2
    (2, 800, 72, 350),    # 12 warps x 2
    (2, 1000, 30, 450),   # 16 warps x 2
    (4, 300, 72, 120),    # 4 warps x 2
    (3, 700, 31, 300),    # 10 warps x 2, odd V (scalar path)
    (2, 1100, 500, 520),  # 16 warps x 4
    (3, 64, 5000, 20),    # CTA-per-row softmax
    (2, 40, 9000, 8),     # big-row fallback
    (4, 30, 2, 3),        # V = 2: only blank + one label
    (16, 10, 6, 4),
])
def test_random_vs_oracle(cuda_dev, B, T, V, Lmax):
    rng = np.random.RandomState(B * 1000 + T + V)
    x, lab, sl = _random_case(rng, B, T, V, Lmax)
    loss, grad = run_gpu(x, lab, sl)
    oloss, ograd = oracle.ctc_loss_grad(x.astype(np.float64), lab, sl)
    assert_close(loss, grad, oloss, ograd, (B, T, V, Lmax))


def test_long_sequence_precision(cuda_dev):
    """T=3000: plain fp32 log-space drifts ~1e-3 here; the fp64-state / fp32-transcendental lattice must not."""
    rng = np.random.RandomState(1)
    x, lab, sl = _random_case(rng, 2, 3000, 72, 300, full_len=True)
    loss, grad = run_gpu(x, lab, sl)
    oloss, ograd = oracle.ctc_loss_grad(x.astype(np.float64), lab, sl)
    assert_close(loss, grad, oloss, ograd, "T3000")


def test_invalid_label_raises(cuda_dev):
    from lstm_ctc_b200 import _lib
    x = np.zeros((2, 6, 5), dtype=np.float32)
    with pytest.raises(_lib.InvalidArgumentError):
        run_gpu(x, [[1, 4], [0, -1]], [6, 6])      # 4 == blank id


def test_gradient_rows_sum_to_zero_at_full_size(cuda_dev):
    """Size-independent property at a BASELINE sweep size (B=256): for every live frame
    sum_v grad = 0 (softmax sums to 1, occupancies sum to 1); padded frames are exactly 0; and the
    loss equals -log p recomputed from a linear functional of the gradient at t=0."""
    from lstm_ctc_b200.ctc import ctc_loss_grad
    d = torch.device("cuda:0")
    g = torch.Generator(device="cpu").manual_seed(3)
    B, T, V, L = 256, 700, 72, 80
    x = (torch.randn(B, T, V, generator=g) * 3).to(d)
    sl = torch.randint(int(0.8 * T), T + 1, (B,), generator=g).to(torch.int32)
    lab = torch.randint(0, V - 1, (B, L), generator=g)
    loss, grad = ctc_loss_grad(x, lab.to(d), sl.to(d), lattice_layout=_LAYOUT[0])
    rows = grad.sum(-1).cpu()
    assert rows.abs().max() < 2e-5
    mask = torch.arange(T).unsqueeze(0) >= sl.unsqueeze(1)
    assert grad.cpu()[mask].abs().max() == 0
    assert torch.isfinite(loss).all() and (loss > 0).all()
    # spot-check 4 utterances against the oracle
    idx = [0, 77, 128, 255]
    ol, og = oracle.ctc_loss_grad(x[idx].cpu().numpy().astype(np.float64), lab[idx].numpy(), sl[idx].numpy())
    assert_close(loss[idx].cpu().numpy().astype(np.float64), grad[idx].cpu().numpy().astype(np.float64), ol, og, "B256")


def test_autograd_wrapper(cuda_dev):
    from lstm_ctc_b200.ctc import ctc_loss
    d = torch.device("cuda:0")
    rng = np.random.RandomState(4)
    x, lab, sl = _random_case(rng, 4, 20, 9, 5)
    xt = torch.tensor(x, device=d, requires_grad=True)
    loss = ctc_loss(torch.tensor(lab, device=d), xt, torch.tensor(sl, dtype=torch.int32, device=d))
    (loss * torch.tensor([1.0, 2.0, 0.5, 1.0], device=d)).sum().backward()
    _, og = oracle.ctc_loss_grad(x.astype(np.float64), lab, sl)
    og *= np.array([1.0, 2.0, 0.5, 1.0])[:, None, None]
    assert np.abs(xt.grad.cpu().numpy() - og).max() < 2e-4


def test_default_layout_by_batch_size(cuda_dev):
    """The plain entry point picks the layout from the batch size; both choices agree with the forced layouts up to the order
    of the gradient's atomic adds (B = 4: two CTAs per utterance; B = 80 > SMs / 2: one)."""
    from lstm_ctc_b200.ctc import ctc_loss_grad
    d = torch.device("cuda:0")
    for B in (4, 80):
        rng = np.random.RandomState(B)
        x, lab, sl = _random_case(rng, B, 90, 72, 30)
        a = [torch.tensor(x, device=d), torch.tensor(lab, device=d), torch.tensor(sl, dtype=torch.int32, device=d)]
        l_def, g_def = ctc_loss_grad(*a)
        for lay in (0, 1):
            l, g = ctc_loss_grad(*a, lattice_layout=lay)
            fin = torch.isfinite(l_def)
            assert torch.equal(torch.isfinite(l), fin)
            assert torch.allclose(l[fin], l_def[fin], rtol=1e-6, atol=1e-6) and (g - g_def).abs().max() < 2e-6


@pytest.mark.parametrize("B,T,V,Lmax,nchk", [(2, 2300, 30, 1100, 2), (80, 2300, 30, 1100, 2), (2, 4300, 30, 2100, 1)])
def test_very_long_label_sequences(cuda_dev, B, T, V, Lmax, nchk):
    """More than 1023 labels per utterance: 8 / 16 lattice states per thread (always the two-CTA layout, whatever the batch size or
    the requested layout).  Full-length labels so that the coarse mappings are really the ones running."""
    rng = np.random.RandomState(T + B)
    x = (rng.randn(B, T, V) * 3).astype(np.float32)
    sl = np.full(B, T); sl[-1] = T - 3
    lab = rng.randint(0, V - 1, size=(B, Lmax)).astype(np.int64)
    loss, grad = run_gpu(x, lab, sl)
    idx = list(range(B))[:nchk - 1] + [B - 1]
    oloss, ograd = oracle.ctc_loss_grad(x[idx].astype(np.float64), lab[idx], sl[idx])
    assert_close(loss[idx], grad[idx], oloss, ograd, (B, T, V, Lmax))
    assert np.isfinite(loss).all() and np.abs(grad.sum(-1)).max() < 5e-5
