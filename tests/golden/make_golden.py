"""Generates tests/golden/*.json.  Run from the repo root:  python tests/golden/make_golden.py

ctc_tf_kat.json : the two known-answer vectors of upstream TensorFlow's
                  tensorflow/python/kernel_tests/ctc_loss_op_test.py (testBasic), the only external
                  pins available for tf.nn.ctc_loss semantics (the reference tree has no tests).
                  Expected losses/gradient rows are the published constants; the generating script
                  re-derives them with torch.nn.functional.ctc_loss and refuses to write on mismatch.
ctc_cases.json  : small seeded cases (ragged lengths, repeats, empty labels, too-long labels,
                  infeasible repeats) with outputs of oracle/ctc_oracle.c, cross-checked against
                  torch.nn.functional.ctc_loss where torch defines the case.
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))

P0 = [[0.633766, 0.221185, 0.0917319, 0.0129757, 0.0142857, 0.0260553],
      [0.111121, 0.588392, 0.278779, 0.0055756, 0.00569609, 0.010436],
      [0.0357786, 0.633813, 0.321418, 0.00249248, 0.00272882, 0.0037688],
      [0.0663296, 0.643849, 0.280111, 0.00283995, 0.0035545, 0.00331533],
      [0.458235, 0.396634, 0.123377, 0.00648837, 0.00903441, 0.00623107]]
P1 = [[0.30176, 0.28562, 0.0831517, 0.0862751, 0.0816851, 0.161508],
      [0.24082, 0.397533, 0.0557226, 0.0546814, 0.0557528, 0.19549],
      [0.230246, 0.450868, 0.0389607, 0.038309, 0.0391602, 0.202456],
      [0.280884, 0.429522, 0.0326593, 0.0339046, 0.0326856, 0.190345],
      [0.423286, 0.315517, 0.0338439, 0.0393744, 0.0339315, 0.154046]]
G0 = [[-0.366234, 0.221185, 0.0917319, 0.0129757, 0.0142857, 0.0260553],
      [0.111121, -0.411608, 0.278779, 0.0055756, 0.00569609, 0.010436],
      [0.0357786, 0.633813, -0.678582, 0.00249248, 0.00272882, 0.0037688],
      [0.0663296, -0.356151, 0.280111, 0.00283995, 0.0035545, 0.00331533],
      [-0.541765, 0.396634, 0.123377, 0.00648837, 0.00903441, 0.00623107]]
G1 = [[-0.69824, 0.28562, 0.0831517, 0.0862751, 0.0816851, 0.161508],
      [0.24082, -0.602467, 0.0557226, 0.0546814, 0.0557528, 0.19549],
      [0.230246, 0.450868, 0.0389607, 0.038309, 0.0391602, -0.797544],
      [0.280884, -0.570478, 0.0326593, 0.0339046, 0.0326856, 0.190345],
      [-0.576714, 0.315517, 0.0338439, 0.0393744, 0.0339315, 0.154046]]


def torch_ctc(logits, labels, seq_len):
    x = torch.tensor(logits, dtype=torch.float64, requires_grad=True)
    lab = torch.tensor(labels)
    tl = (lab >= 0).sum(1)
    lp = torch.log_softmax(x, -1).transpose(0, 1)
    loss = torch.nn.functional.ctc_loss(lp, lab.clamp(min=0), torch.tensor(seq_len).long(), tl,
                                        blank=x.shape[-1] - 1, reduction="none")
    loss.sum().backward()
    return loss.detach().numpy(), x.grad.numpy()


def main():
    kat = {"source": "tensorflow/python/kernel_tests/ctc_loss_op_test.py::testBasic (TF r1.8), blank = 5",
           "probs": [P0, P1], "labels": [[0, 1, 2, 1, 0], [0, 1, 1, 0, -1]], "seq_len": [5, 5],
           "loss": [3.34211, 5.42262], "grad": [G0, G1]}
    lg = np.log(np.array(kat["probs"]))
    tl, tg = torch_ctc(lg, kat["labels"], kat["seq_len"])
    assert np.allclose(tl, kat["loss"], atol=2e-5), tl
    assert np.allclose(tg, np.array(kat["grad"]), atol=2e-6), np.abs(tg - np.array(kat["grad"])).max()
    ol, og = oracle.ctc_loss_grad(lg, np.array(kat["labels"]), np.array(kat["seq_len"]))
    assert np.allclose(ol, kat["loss"], atol=2e-5) and np.allclose(og, np.array(kat["grad"]), atol=2e-6)
    json.dump(kat, open(os.path.join(HERE, "ctc_tf_kat.json"), "w"), indent=1)

    rng = np.random.RandomState(777)
    cases = []
    specs = [  # (B, T, V, Lmax, description)
        (4, 12, 7, 4, "ragged"), (3, 9, 5, 0, "empty_labels"), (3, 6, 4, 6, "labels_longer_than_input"),
        (2, 5, 4, 3, "infeasible_repeats"), (5, 40, 30, 12, "wider"), (2, 33, 72, 10, "v72"),
    ]
    for (B, T, V, Lmax, desc) in specs:
        logits = rng.randn(B, T, V) * 2.0
        seq_len = rng.randint(max(1, T // 2), T + 1, size=B)
        labels = -np.ones((B, max(Lmax, 1)), dtype=np.int64)
        for b in range(B):
            L = 0 if Lmax == 0 else rng.randint(1, Lmax + 1)
            if desc == "labels_longer_than_input":
                L = Lmax if b == 0 else min(L, int(seq_len[b]) // 2)
                if b == 0:
                    seq_len[b] = Lmax - 1
            labels[b, :L] = rng.randint(0, V - 1, size=L)
        if desc == "infeasible_repeats":
            labels[0, :3] = [1, 1, 1]
            seq_len[0] = 4               # needs 5 frames -> no valid path
            labels[1, :3] = [0, 1, 0]
            seq_len[1] = 5
        if desc == "ragged":
            seq_len[1] = 0               # zero-length utterance -> skipped
        loss, grad = oracle.ctc_loss_grad(logits, labels, seq_len)
        cases.append({"desc": desc, "logits": np.round(logits, 6).tolist(), "labels": labels.tolist(),
                      "seq_len": seq_len.tolist(),
                      "loss": [float(v) if np.isfinite(v) else "inf" for v in loss],
                      "grad": np.round(grad, 9).tolist()})
    # recompute with the rounded logits so the stored outputs match the stored inputs exactly
    for c in cases:
        lg = np.array(c["logits"])
        loss, grad = oracle.ctc_loss_grad(lg, np.array(c["labels"]), np.array(c["seq_len"]))
        c["loss"] = [float(v) if np.isfinite(v) else "inf" for v in loss]
        c["grad"] = np.round(grad, 9).tolist()
        # torch cross-check on the utterances torch defines identically (feasible, non-skipped)
        tl, tg = torch_ctc(lg, c["labels"], c["seq_len"])
        for b in range(len(loss)):
            L = sum(1 for v in c["labels"][b] if v >= 0)
            if c["seq_len"][b] > 0 and L <= c["seq_len"][b] and np.isfinite(loss[b]):
                assert abs(tl[b] - loss[b]) < 1e-8 * max(1, abs(loss[b])), (c["desc"], b, tl[b], loss[b])
                assert np.abs(tg[b] - grad[b]).max() < 1e-8, (c["desc"], b)
    json.dump(cases, open(os.path.join(HERE, "ctc_cases.json"), "w"))
    print("wrote", len(cases), "cases")


if __name__ == "__main__":
    main()
