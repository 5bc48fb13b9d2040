"""CPU: pins the CTC oracle (oracle/ctc_oracle.c) against the TF known-answer vectors, the committed
golden cases, torch.nn.functional.ctc_loss and brute-force path enumeration."""
import json
import os

import numpy as np
import pytest
import torch

import oracle
from oracle.ctc import ctc_brute_force, ctc_loss_numpy

G = os.path.join(os.path.dirname(__file__), "golden")


def test_tf_known_answers():
    kat = json.load(open(os.path.join(G, "ctc_tf_kat.json")))
    loss, grad = oracle.ctc_loss_grad(np.log(np.array(kat["probs"])), np.array(kat["labels"]), np.array(kat["seq_len"]))
    assert np.allclose(loss, kat["loss"], atol=2e-5)
    assert np.allclose(grad, np.array(kat["grad"]), atol=2e-6)


def test_golden_cases_roundtrip():
    cases = json.load(open(os.path.join(G, "ctc_cases.json")))
    assert len(cases) >= 6
    for c in cases:
        loss, grad = oracle.ctc_loss_grad(np.array(c["logits"]), np.array(c["labels"]), np.array(c["seq_len"]))
        exp = np.array([np.inf if v == "inf" else v for v in c["loss"]])
        assert np.allclose(loss, exp, rtol=1e-9, atol=1e-9), c["desc"]
        assert np.allclose(grad, np.array(c["grad"]), atol=2e-9), c["desc"]


def test_edge_semantics():
    cases = {c["desc"]: c for c in json.load(open(os.path.join(G, "ctc_cases.json")))}
    c = cases["labels_longer_than_input"]          # ignore_longer_outputs_than_inputs=True -> 0 / 0
    assert c["loss"][0] == 0.0 and np.abs(np.array(c["grad"][0])).max() == 0.0
    c = cases["ragged"]                              # seq_len == 0 -> skipped
    assert c["loss"][1] == 0.0 and np.abs(np.array(c["grad"][1])).max() == 0.0
    c = cases["infeasible_repeats"]                  # no valid path -> +inf, grad = softmax
    assert c["loss"][0] == "inf"
    g = np.array(c["grad"][0]); x = np.array(c["logits"][0]); T = c["seq_len"][0]
    sm = np.exp(x[:T] - x[:T].max(1, keepdims=True)); sm /= sm.sum(1, keepdims=True)
    assert np.allclose(g[:T], sm, atol=1e-8) and np.abs(g[T:]).max() == 0
    c = cases["empty_labels"]                        # L = 0: loss = -sum log y(blank)
    x = np.array(c["logits"][0]); T = c["seq_len"][0]
    lp = x[:T] - np.log(np.exp(x[:T]).sum(1, keepdims=True))
    assert abs(c["loss"][0] + lp[:, -1].sum()) < 1e-8


def test_bad_label_raises():
    with pytest.raises(ValueError):
        oracle.ctc_loss_grad(np.zeros((1, 4, 5)), np.array([[4]]), np.array([4]))   # blank id as label


@pytest.mark.parametrize("seed", range(3))
def test_vs_torch_random(seed):
    rng = np.random.RandomState(seed)
    B, T, V, L = 6, 25, 11, 7
    x = rng.randn(B, T, V) * 3
    sl = rng.randint(15, T + 1, size=B)
    lab = -np.ones((B, L), dtype=np.int64)
    for b in range(B):
        n = rng.randint(1, L + 1)
        lab[b, :n] = rng.randint(0, V - 1, size=n)
    xt = torch.tensor(x, requires_grad=True)
    lp = torch.log_softmax(xt, -1).transpose(0, 1)
    tl = torch.nn.functional.ctc_loss(lp, torch.tensor(lab).clamp(min=0), torch.tensor(sl).long(),
                                      torch.tensor((lab >= 0).sum(1)), blank=V - 1, reduction="none")
    tl.sum().backward()
    ol, og = oracle.ctc_loss_grad(x, lab, sl)
    assert np.allclose(ol, tl.detach().numpy(), rtol=1e-10)
    assert np.abs(og - xt.grad.numpy()).max() < 1e-9


def test_vs_numpy_twin_and_bruteforce():
    rng = np.random.RandomState(5)
    T, V = 5, 4
    x = rng.randn(1, T, V)
    for label in ([0], [1, 1], [0, 2, 1], []):
        lab = -np.ones((1, 3), dtype=np.int64)
        lab[0, :len(label)] = label
        ol, _ = oracle.ctc_loss_grad(x, lab, np.array([T]))
        assert abs(ol[0] - ctc_loss_numpy(x[0], label, T)) < 1e-10
        assert abs(ol[0] - ctc_brute_force(x[0], label, T)) < 1e-10


def test_gradient_is_derivative():
    rng = np.random.RandomState(9)
    x = rng.randn(2, 8, 6)
    lab = np.array([[0, 1, 1], [3, -1, -1]])
    sl = np.array([8, 6])
    _, g = oracle.ctc_loss_grad(x, lab, sl)
    eps = 1e-6
    for (b, t, v) in [(0, 0, 0), (0, 3, 5), (1, 5, 3), (1, 7, 0)]:
        xp = x.copy(); xp[b, t, v] += eps
        xm = x.copy(); xm[b, t, v] -= eps
        num = (oracle.ctc_loss_grad(xp, lab, sl)[0].sum() - oracle.ctc_loss_grad(xm, lab, sl)[0].sum()) / (2 * eps)
        assert abs(num - g[b, t, v]) < 1e-6
