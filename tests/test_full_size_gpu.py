"""Size-independent properties of the whole hot path AT BASELINE's full sizes (C3: 5 x BiLSTM 512/dir, K=8 mixture,
64 utterances x ~1500 frames), where the fp64 oracle cannot run in seconds:

  * utterances are independent through BiLSTM, mixture layer and CTC, so loss and parameter gradients of
    the 64-utterance batch equal the SUM over its two 32-utterance halves (linearity of the summed loss, graph.py:116)
    -- which also exercises different cluster groupings, time lengths and the split-launch recurrence;
  * length masking is exact: encoder rows past sequence_length are exactly 0, and their logits equal the output layer
    applied to a zero row (Q5);
  * dlogits rows sum to 0 on live frames and are exactly 0 on padded frames.
Dropout is off (keep = 1) so the halves see the same function."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_c3_batch_equals_sum_of_halves_and_masking(cuda_dev):
    import bench
    from lstm_ctc_b200.model import AcousticModel
    w = bench.WORKLOADS["c3"]
    cfg = bench.nnet_config(w, keep=1.0)
    m = AcousticModel(cfg, cuda_dev, seed=99)
    x, lens, y = [t.to(cuda_dev) for t in bench.synth_batch(w, 4242)]
    B, T = x.shape[0], x.shape[1]

    loss_all, per_utt = m.loss_and_grad(x, lens, y, check_labels=True)
    g_all = m.params.gflat.clone()
    logits = m._out_ws(T, B)["logits"].clone()
    enc_rows = m._top[0].float().view(T, B, -1).clone()
    torch.cuda.synchronize()
    assert torch.isfinite(per_utt).all() and torch.isfinite(g_all).all()

    # ---- exact masking at full size ----
    pad = (torch.arange(T, device=cuda_dev).unsqueeze(1) >= lens.unsqueeze(0))          # [T, B]
    assert pad.any()
    assert enc_rows[pad].abs().max().item() == 0.0
    zero_row_logits = logits[0, T - 1] if int(lens[0]) < T else None
    pl = logits.permute(1, 0, 2)[pad]                                                     # logits of all padded frames
    assert (pl - pl[0]).abs().max().item() < 1e-5                                         # all equal: MoE(0) / bias only
    del zero_row_logits

    # ---- bit-reproducibility: same batch again, and (below) the halves -- each MMA issuer thread of the recurrence owns its
    # accumulator, so the activations do not depend on thread interleaving, on the batch an utterance travels in, or on
    # whether the recurrence ran as one launch or two.  (A randomly initialised 5-layer peephole LSTM amplifies 1e-7
    # perturbations to O(1) within ~800 frames, so anything less than bit-equality would make this test meaningless.)
    enc_bits = m._top[0].view(T, B, -1).clone()
    m.loss_and_grad(x, lens, y, check_labels=False)
    assert torch.equal(m._top[0].view(T, B, -1), enc_bits)

    # ---- halves ----
    g_sum = torch.zeros_like(g_all)
    loss_sum = 0.0
    per = []
    for sl in (slice(0, B // 2), slice(B // 2, B)):
        Th = int(lens[sl].max())
        ls, pu = m.loss_and_grad(x[sl, :Th].contiguous(), lens[sl].contiguous(), y[sl].contiguous(), check_labels=False)
        assert torch.equal(m._top[0].view(Th, B // 2, -1), enc_bits[:Th, sl])        # encoder output: bit-identical per utterance
        g_sum += m.params.gflat
        loss_sum += float(ls)
        per.append(pu.clone())
    torch.cuda.synchronize()
    per = torch.cat(per)
    assert abs(loss_sum - float(loss_all)) < 2e-4 * abs(float(loss_all))
    assert (per - per_utt).abs().max().item() < 2e-3 * per_utt.abs().max().item()
    # gradients: normwise per variable (fp16/bf16 operand rounding differs only through accumulation order / grouping)
    for name in m.params.order:
        s = m.params.specs[name]
        a, b = g_all[s.offset:s.offset + s.numel], g_sum[s.offset:s.offset + s.numel]
        rel = ((a - b).norm() / (a.norm() + 1e-20)).item()
        assert rel < 2e-2, (name, rel)
    from lstm_ctc_b200 import _lib
    assert _lib.lib().lcb_device_error(1) == 0


def test_c3_ctc_gradient_rows(cuda_dev):
    import bench
    from lstm_ctc_b200.ctc import ctc_loss_grad
    w = bench.WORKLOADS["c3"]
    g = torch.Generator().manual_seed(5)
    B, T, V = w["B"], w["T"], w["V"]
    _, lens, y = bench.synth_batch(w, 31)
    logits = (torch.randn(B, T, V, generator=g) * 3).to(cuda_dev)
    loss, grad = ctc_loss_grad(logits, y.to(cuda_dev), lens.to(cuda_dev))
    rows = grad.sum(-1)
    pad = torch.arange(T, device=cuda_dev).unsqueeze(0) >= lens.to(cuda_dev).unsqueeze(1)
    assert rows[~pad].abs().max().item() < 2e-5
    assert grad[pad].abs().max().item() == 0.0
    assert torch.isfinite(loss).all() and (loss > 0).all()
