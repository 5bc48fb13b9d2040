"""GPU parity of nnet_type 'lstm' -- the uni-directional residual stack (functional core of nnet/lstm.py:125-368) run on the
BiLSTM kernels with zero backward cells -- vs the fp64 oracle (oracle.lstm_*).  Tolerances as for the BiLSTM path: logits 1e-2 of
max|logit|, summed loss 2e-3, per-variable gradients 5e-2 normwise; the zero half must stay EXACTLY zero under training."""
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

CASES = {
    "affine_d24": dict(input_dim=24, num_layers=3, num_neurons=64, num_projects=32, num_targets=12, use_peepholes=True, num_experts=0),
    "mos_k4_residual0": dict(input_dim=64, num_layers=2, num_neurons=128, num_projects=64, num_targets=13, use_peepholes=True, num_experts=4),
}


def make_batch(cfg, B, T, Lmax, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, cfg.input_dim, generator=g, dtype=torch.float64)
    lens = torch.randint(max(2, int(0.7 * T)), T + 1, (B,), generator=g).to(torch.int32)
    lens[0] = T
    labels = -torch.ones(B, Lmax, dtype=torch.int64)
    for b in range(B):
        x[b, lens[b]:] = 0
        n = int(torch.randint(1, min(Lmax, int(lens[b]) // 2) + 1, (1,), generator=g))
        labels[b, :n] = torch.randint(0, cfg.num_targets - 1, (n,), generator=g)
    return x, lens, labels


def nnet_config(cfg, keep=1.0):
    return {"nnet_type": "lstm", "input_dim": cfg.input_dim, "left_context": 0, "right_context": 0,
            "num_layers": cfg.num_layers, "num_neurons": cfg.num_neurons, "num_projects": cfg.num_projects,
            "num_targets": cfg.num_targets, "num_experts": cfg.num_experts, "moe_temp": cfg.moe_temp, "dropout_rate": keep}


def _mask(n, keep, seed, dev):
    from lstm_ctc_b200 import _lib
    m = torch.empty(n, dtype=torch.uint8, device=dev)
    _lib.check(_lib.lib().lcb_dropout_mask(_lib.ptr(m), n, keep, seed, _lib.stream_ptr()), "mask")
    return m.cpu().double()


def _compare(m, p64, ctc, ref_logits, loss_sum, lens, T, B):
    logits = m._out_ws(T, B)["logits"].cpu().double()
    live = (torch.arange(T).unsqueeze(0) < lens.unsqueeze(1))
    err = (logits - ref_logits.detach())[live].abs().max().item()
    assert err < 1e-2 * ref_logits.abs().max().item(), ("logits", err)
    assert abs(loss_sum.item() - ctc.item()) < 2e-3 * abs(ctc.item()), (loss_sum.item(), ctc.item())
    grads = {k: v.cpu().double() for k, v in m.to_tf_dict(grads=True).items()}
    assert set(grads) == set(p64)
    bad = {k: ((v - p64[k].grad).norm() / (p64[k].grad.norm() + 1e-12)).item() for k, v in grads.items()}
    bad = {k: r for k, r in bad.items() if r > 5e-2}
    assert not bad, bad


@pytest.mark.parametrize("name", list(CASES))
def test_logits_loss_grads_vs_oracle(cuda_dev, name):
    from lstm_ctc_b200 import _lib
    from lstm_ctc_b200.lstm import backward_half_is_zero
    from lstm_ctc_b200.model import AcousticModel
    cfg = oracle.OracleConfig(**CASES[name])
    params = oracle.init_lstm_params(cfg, seed=11, bias_scale=0.1)
    B, T = 6, 30
    x, lens, labels = make_batch(cfg, B=B, T=T, Lmax=8, seed=12)
    p64 = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    ctc, _, ref_logits = oracle.lstm_training_loss(p64, cfg, x, lens, labels, l2_decay_weight=0.0)
    ctc.backward()
    m = AcousticModel(nnet_config(cfg), cuda_dev, init=False)
    m.from_tf_dict(params)
    rt = m.to_tf_dict()
    for k, v in params.items():
        assert torch.equal(rt[k].cpu().double(), v.float().double()), k
    loss_sum, _ = m.loss_and_grad(x.float().to(cuda_dev), lens.to(cuda_dev), labels.to(cuda_dev))
    _compare(m, p64, ctc, ref_logits, loss_sum, lens, T, B)
    # the embedded backward half: zero weights, and EXACTLY zero gradients
    gfull = {k: v.cpu() for k, v in m.to_tf_dict(grads=True, embedded=True).items()}
    assert backward_half_is_zero(m.cfg, gfull)
    assert _lib.lib().lcb_device_error(1) == 0


def test_dropout_parity_with_exported_masks(cuda_dev):
    """keep 0.8: out = dropout(x + cell(x)) -- the residual shares the layer's output mask, in forward and backward."""
    from lstm_ctc_b200.model import AcousticModel
    cfg = oracle.OracleConfig(**CASES["affine_d24"])
    params = oracle.init_lstm_params(cfg, seed=31, bias_scale=0.1)
    B, T, keep = 5, 14, 0.8
    x, lens, labels = make_batch(cfg, B=B, T=T, Lmax=4, seed=32)
    m = AcousticModel(nnet_config(cfg, keep), cuda_dev, init=False)
    m.from_tf_dict(params)
    loss_sum, _ = m.loss_and_grad(x.float().to(cuda_dev), lens.to(cuda_dev), labels.to(cuda_dev))
    P, N = cfg.num_projects, T * B
    masks = {i: _mask(N * 2 * P, keep, m.enc.dropout_seed(i), cuda_dev).view(T, B, 2 * P).permute(1, 0, 2)[:, :, :P].contiguous()
             for i in range(cfg.num_layers)}
    p64 = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    ctc, _, ref_logits = oracle.lstm_training_loss(p64, cfg, x, lens, labels, l2_decay_weight=0.0, keep_prob=keep, masks=masks)
    ctc.backward()
    _compare(m, p64, ctc, ref_logits, loss_sum, lens, T, B)


@pytest.mark.parametrize("opt", ["adam", "momentum"])
def test_training_keeps_the_backward_half_at_zero(cuda_dev, opt):
    """Four optimizer steps (L2, global-norm clipping, Adam / Momentum): the loss sequence follows the fp64 oracle and every
    backward-cell variable, and every weight that reads the backward half, is still exactly 0.0 afterwards."""
    from lstm_ctc_b200.lstm import backward_half_is_zero
    from lstm_ctc_b200.model import AcousticModel
    cfg = oracle.OracleConfig(**CASES["mos_k4_residual0"])
    params = oracle.init_lstm_params(cfg, seed=41, bias_scale=0.05)
    x, lens, labels = make_batch(cfg, B=4, T=20, Lmax=5, seed=42)
    m = AcousticModel(nnet_config(cfg, 0.9), cuda_dev, init=False)      # dropout on: the masks must not disturb the zeros either
    m.from_tf_dict(params)
    losses = []
    for step in range(4):
        loss_sum, _ = m.loss_and_grad(x.float().to(cuda_dev), lens.to(cuda_dev), labels.to(cuda_dev))
        m.optimizer_step(opt, 1e-3, clip_norm=5.0, l2_decay_weight=1e-5)
        losses.append(loss_sum.item())
    full = {k: v.cpu() for k, v in m.to_tf_dict(embedded=True).items()}
    assert backward_half_is_zero(m.cfg, full)
    moved = sum((m.to_tf_dict()[k].cpu().double() - params[k].float().double()).abs().max().item() for k in params)
    assert moved > 0 and all(torch.isfinite(torch.tensor(losses)))
    # without dropout the loss sequence follows the oracle
    m2 = AcousticModel(nnet_config(cfg, 1.0), cuda_dev, init=False)
    m2.from_tf_dict(params)
    p = {k: v.clone() for k, v in params.items()}
    state = {}
    for step in range(3):
        pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
        ctc, total, _ = oracle.lstm_training_loss(pr, cfg, x, lens, labels, l2_decay_weight=1e-5)
        total.backward()
        clipped, _ = oracle.clip_by_global_norm({k: v.grad for k, v in pr.items()}, 5.0)
        stepf = oracle.adam_step if opt == "adam" else oracle.momentum_step
        p = stepf({k: v.detach() for k, v in pr.items()}, clipped, state, 1e-3)
        loss_sum, _ = m2.loss_and_grad(x.float().to(cuda_dev), lens.to(cuda_dev), labels.to(cuda_dev))
        m2.optimizer_step(opt, 1e-3, clip_norm=5.0, l2_decay_weight=1e-5)
        assert abs(loss_sum.item() - ctc.item()) < 3e-3 * abs(ctc.item()), (step, loss_sum.item(), ctc.item())
    assert backward_half_is_zero(m2.cfg, {k: v.cpu() for k, v in m2.to_tf_dict(embedded=True).items()})


def test_graph_api_accepts_nnet_type_lstm(cuda_dev):
    """create_graph_for_training_ctc / Session.run with nnet_type 'lstm': a few steps through the reference-facing API."""
    import numpy as np
    import lstm_ctc_b200 as nnet
    cfg = oracle.OracleConfig(**CASES["affine_d24"])
    x, lens, labels = make_batch(cfg, B=4, T=16, Lmax=4, seed=52)

    class DS:
        def __iter__(self):
            for _ in range(3):
                for b in range(4):
                    n = int(lens[b])
                    yield {"nnet_input": x[b, :n].float().numpy(), "nnet_target": labels[b][labels[b] >= 0].numpy()}

    init, pipeline = nnet.create_pipeline_sequence_batch(DS(), cfg.input_dim, batch_size=4)
    graph = nnet.create_graph_for_training_ctc(pipeline, nnet_config(cfg), learn_rate=1e-3, optimizer="adam", seed=3)
    sess = nnet.Session()
    sess.run(init)
    vals = [sess.run({k: graph[k] for k in ("train", "eval_loss", "size")}) for _ in range(3)]
    assert all(np.isfinite(v["eval_loss"]) for v in vals) and vals[-1]["eval_loss"] < vals[0]["eval_loss"]
    sd = nnet.trainable_variables().state_dict()                       # what a checkpoint holds: the uni-directional variables
    assert sorted(sd) == sorted(oracle.lstm_param_order(cfg))
    assert nnet.get_create_logits("lstm") is not None and nnet.get_create_logits("cudnnlstm") is None
