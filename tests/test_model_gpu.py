"""GPU parity of the whole hot path (BiLSTM -> mixture/affine output -> CTC -> gradients -> L2/clip/optimizer)
vs the fp64 oracle (oracle/model.py), same synthetic inputs and weights.

Stated tolerances: logits within 1e-2 * max|logit| (fp16 operands / fp32 accumulate); summed CTC loss within
2e-3 relative; per-variable gradients within 5e-2 normwise relative; the optimizer kernel itself (fp32 maths on
given gradients) within 1e-5."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


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


def nnet_config(cfg):
    return {"nnet_type": "blstm", "input_dim": cfg.input_dim, "left_context": 0, "right_context": 0,
            "num_layers": cfg.num_layers, "num_neurons": cfg.num_neurons, "num_projects": cfg.num_projects,
            "num_targets": cfg.num_targets, "use_peepholes": cfg.use_peepholes, "num_experts": cfg.num_experts,
            "moe_temp": cfg.moe_temp, "dropout_rate": 1.0}


CASES = {
    "affine": dict(input_dim=40, num_layers=2, num_neurons=128, num_projects=64, num_targets=20, use_peepholes=True, num_experts=0),
    "mos_k4": dict(input_dim=24, num_layers=1, num_neurons=64, num_projects=64, num_targets=13, use_peepholes=False, num_experts=4),
    "mos_k8_v72": dict(input_dim=40, num_layers=2, num_neurons=128, num_projects=128, num_targets=72, use_peepholes=True, num_experts=8),
    "mos_k12": dict(input_dim=24, num_layers=1, num_neurons=64, num_projects=64, num_targets=11, use_peepholes=True, num_experts=12),
    "mos_k5_odd": dict(input_dim=16, num_layers=1, num_neurons=64, num_projects=32, num_targets=31, use_peepholes=True, num_experts=5),
}


@pytest.mark.parametrize("name", list(CASES))
def test_logits_loss_grads_vs_oracle(cuda_dev, name):
    from lstm_ctc_b200.model import AcousticModel
    cfg = oracle.OracleConfig(**CASES[name])
    params = oracle.init_params(cfg, seed=11, bias_scale=0.1)
    x, lens, labels = make_batch(cfg, B=6, T=30, Lmax=8, seed=12)
    p64 = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    ctc, total, ref_logits = oracle.training_loss(p64, cfg, x, lens, labels, l2_decay_weight=0.0)
    ctc.backward()

    m = AcousticModel(nnet_config(cfg), cuda_dev, init=False)
    m.from_tf_dict(params)
    rt = m.to_tf_dict()
    for k, v in params.items():
        assert torch.equal(rt[k].cpu().double(), v.float().double()), k                  # exact layout round trip
    loss_sum, loss = m.loss_and_grad(x.float().to(cuda_dev), lens.to(cuda_dev), labels.to(cuda_dev))
    logits = m._out_ws(30, 6)["logits"].cpu().double()
    scale = ref_logits.abs().max().item()
    mask = (torch.arange(30).unsqueeze(0) < lens.unsqueeze(1))
    err = (logits - ref_logits.detach())[mask].abs().max().item()
    assert err < 1e-2 * scale, ("logits", err, scale)
    assert abs(loss_sum.item() - ctc.item()) < 2e-3 * abs(ctc.item()), (loss_sum.item(), ctc.item())
    grads = {k: v.cpu().double() for k, v in m.to_tf_dict(grads=True).items()}
    bad = {}
    for k, v in grads.items():
        rg = p64[k].grad
        rel = ((v - rg).norm() / (rg.norm() + 1e-12)).item()
        if rel > 5e-2:
            bad[k] = rel
    assert not bad, bad
    from lstm_ctc_b200 import _lib
    assert _lib.lib().lcb_device_error(1) == 0


@pytest.mark.parametrize("opt", ["sgd", "momentum", "adam"])
def test_optimizer_kernel_vs_oracle(cuda_dev, opt):
    """L2 (skipping 'bias' variables) + clip_by_global_norm + update, 3 consecutive steps, on a tiny model's
    flat buffers with injected gradients."""
    from lstm_ctc_b200.model import AcousticModel
    cfg = oracle.OracleConfig(input_dim=8, num_layers=1, num_neurons=64, num_projects=8, num_targets=5, use_peepholes=True, num_experts=2)
    params = oracle.init_params(cfg, seed=3, bias_scale=0.1)
    m = AcousticModel(nnet_config(cfg), cuda_dev, init=False)
    m.from_tf_dict(params)
    p = {k: v.clone().float().double() for k, v in params.items()}
    state = {}
    g = torch.Generator().manual_seed(0)
    for step in range(3):
        grads = {k: torch.randn(v.shape, generator=g, dtype=torch.float64) * (3.0 if step == 1 else 0.01) for k, v in p.items()}
        # inject: write the TF-layout gradients into the device-layout flat gradient buffer
        m.params.gflat.zero_()
        shadow = AcousticModel(nnet_config(cfg), cuda_dev, init=False)
        shadow.from_tf_dict(grads)
        m.params.gflat.copy_(shadow.params.flat)
        full = {k: grads[k] + (1e-3 * p[k] if "bias" not in k else 0) for k in p}
        clipped, gn = oracle.clip_by_global_norm(full, 5.0)
        if opt == "adam":
            p = oracle.adam_step(p, clipped, state, 1e-2)
        elif opt == "momentum":
            p = oracle.momentum_step(p, clipped, state, 1e-2)
        else:
            p = oracle.sgd_step(p, clipped, state, 1e-2)
        m.optimizer_step(opt, 1e-2, clip_norm=5.0, l2_decay_weight=1e-3)
        assert abs(m.last_grad_norm() - gn) < 1e-4 * gn
        ours = m.to_tf_dict()
        for k in p:
            assert (ours[k].cpu().double() - p[k]).abs().max().item() < 2e-5, (opt, step, k)


def test_training_steps_track_oracle(cuda_dev):
    """Three SGD steps on the same batch: loss sequence and final weights follow the fp64 oracle."""
    from lstm_ctc_b200.model import AcousticModel
    cfg = oracle.OracleConfig(**CASES["mos_k4"])
    params = oracle.init_params(cfg, seed=21, bias_scale=0.05)
    x, lens, labels = make_batch(cfg, B=4, T=20, Lmax=5, seed=22)
    m = AcousticModel(nnet_config(cfg), cuda_dev, init=False)
    m.from_tf_dict(params)
    p = {k: v.clone() for k, v in params.items()}
    for step in range(3):
        pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
        ctc, total, _ = oracle.training_loss(pr, cfg, x, lens, labels, l2_decay_weight=1e-5)
        total.backward()
        clipped, _ = oracle.clip_by_global_norm({k: v.grad for k, v in pr.items()}, 5.0)
        p = oracle.sgd_step({k: v.detach() for k, v in pr.items()}, clipped, {}, 1e-3)
        loss_sum, _ = m.loss_and_grad(x.float().to(cuda_dev), lens.to(cuda_dev), labels.to(cuda_dev))
        m.optimizer_step("sgd", 1e-3, clip_norm=5.0, l2_decay_weight=1e-5)
        assert abs(loss_sum.item() - ctc.item()) < 3e-3 * abs(ctc.item()), (step, loss_sum.item(), ctc.item())
    ours = m.to_tf_dict()
    for k in p:
        d_ref = (p[k] - params[k]).norm().item()
        d_err = (ours[k].cpu().double() - p[k]).norm().item()
        assert d_err < 0.1 * d_ref + 1e-7, (k, d_err, d_ref)


def _mask(n, keep, seed, dev):
    from lstm_ctc_b200 import _lib
    m = torch.empty(n, dtype=torch.uint8, device=dev)
    _lib.check(_lib.lib().lcb_dropout_mask(_lib.ptr(m), n, keep, seed, _lib.stream_ptr()), "mask")
    return m.cpu().double()


@pytest.mark.parametrize("case", ["mos_k4", "mos_k12", "mos_k5_odd"])
def test_dropout_parity_with_exported_masks(cuda_dev, case):
    """keep_prob = 0.8 on LSTM layer outputs (bilstm.py:128,137), mixture weights (moe.py:46) and expert logits
    (moe.py:61).  TF's RNG stream cannot be matched, so the kernels' counter-based masks are exported and fed to
    the oracle: logits, loss and gradients must then agree at the usual tolerance."""
    from lstm_ctc_b200.model import AcousticModel
    cfg = oracle.OracleConfig(**CASES[case])      # K = 4: vectorised mixture backward, experts fixed per lane; 12: atomics; 5: scalar kernel
    cfg.num_layers = 2
    params = oracle.init_params(cfg, seed=31, bias_scale=0.1)
    B, T, keep = 5, 14, 0.8
    x, lens, labels = make_batch(cfg, B=B, T=T, Lmax=4, seed=32)
    nc = nnet_config(cfg); nc["dropout_rate"] = keep
    m = AcousticModel(nc, cuda_dev, init=False)
    m.from_tf_dict(params)
    loss_sum, _ = m.loss_and_grad(x.float().to(cuda_dev), lens.to(cuda_dev), labels.to(cuda_dev))
    P, K, V = cfg.num_projects, cfg.num_experts, cfg.num_targets
    N = T * B
    # export masks (absolute time order, time-major rows n = t*B + b) and convert to the oracle's conventions
    masks = {}
    for i in range(cfg.num_layers):
        mk = _mask(N * 2 * P, keep, m.enc.dropout_seed(i), cuda_dev).view(T, B, 2 * P).permute(1, 0, 2)   # [B,T,2P]
        masks[(i, "f")] = mk[:, :, :P].contiguous()
        masks[(i, "b")] = oracle.model.reverse_sequence(mk[:, :, P:].contiguous(), lens)    # bwd cell runs in reversed time
    seed = m._out_seed
    mp = _mask(N * K, keep, seed, cuda_dev).view(T, B, K).permute(1, 0, 2).reshape(B * T, K, 1)
    md = _mask(N * K * V, keep, seed ^ 0xD1B54A32D192ED03, cuda_dev).view(T, B, V, K).permute(1, 0, 3, 2).reshape(B * T, K, V)
    p64 = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    enc, _ = oracle.blstm_forward(p64, cfg, x, lens, keep_prob=keep, masks=masks)
    y = oracle.create_moe(enc.reshape(-1, 2 * P), p64["Variable"], p64["Variable_1"], p64["Variable_2"], p64["Variable_3"],
                          V, K, cfg.moe_temp, keep_prob=keep, mask_prior=mp, mask_dec=md).reshape(B, T, V)
    ctc = oracle.ctc_loss_sum(y, labels, lens)
    ctc.backward()
    logits = m._out_ws(T, B)["logits"].cpu().double()
    live = (torch.arange(T).unsqueeze(0) < lens.unsqueeze(1))
    assert (logits - y.detach())[live].abs().max().item() < 1e-2 * y.abs().max().item()
    assert abs(loss_sum.item() - ctc.item()) < 2e-3 * abs(ctc.item())
    grads = {k: v.cpu().double() for k, v in m.to_tf_dict(grads=True).items()}
    bad = {k: ((v - p64[k].grad).norm() / (p64[k].grad.norm() + 1e-12)).item() for k, v in grads.items()}
    bad = {k: r for k, r in bad.items() if r > 5e-2}
    assert not bad, bad
    # and: inference mode ignores dropout (bilstm.py:98-99)
    nc2 = dict(nc); nc2["is_training"] = False
    m2 = AcousticModel(nc2, cuda_dev, init=False)
    m2.from_tf_dict(params)
    lg2 = m2.forward_logits(x.float().to(cuda_dev), lens.to(cuda_dev), training=False).cpu().double()
    ref2 = oracle.output_layer(params, cfg, oracle.blstm_forward(params, cfg, x, lens)[0])
    assert (lg2 - ref2)[live].abs().max().item() < 1e-2 * ref2.abs().max().item()


@pytest.mark.parametrize("kind", ["uniform", "prior"])
def test_label_smoothing_regulariser(cuda_dev, kind, tmp_path):
    """reg_loss of bilstm.py:254-269 (KL to uniform / to the class prior, over ALL rows incl. padding) and its
    contribution to the gradients, vs torch autograd on the oracle logits."""
    from lstm_ctc_b200.model import AcousticModel
    cfg = oracle.OracleConfig(**CASES["affine"])
    params = oracle.init_params(cfg, seed=41, bias_scale=0.1)
    B, T = 4, 18
    x, lens, labels = make_batch(cfg, B=B, T=T, Lmax=5, seed=42)
    nc = nnet_config(cfg)
    w = 0.05
    prior = None
    if kind == "uniform":
        nc["uniform_label_sm"] = w
    else:
        counts = np.arange(1, cfg.num_targets + 1, dtype=np.float64)
        p = tmp_path / "label.counts"
        p.write_text("[ " + " ".join(str(c) for c in counts) + " ]\n")
        nc["prior_label_sm"] = w; nc["prior_label_path"] = str(p)
        from lstm_ctc_b200 import get_class_prior
        prior = torch.from_numpy(get_class_prior(str(p))).double()
    m = AcousticModel(nc, cuda_dev, init=False)
    m.from_tf_dict(params)
    loss_sum, _ = m.loss_and_grad(x.float().to(cuda_dev), lens.to(cuda_dev), labels.to(cuda_dev))
    p64 = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    ctc, _, logits = oracle.training_loss(p64, cfg, x, lens, labels, 0.0)
    pr = torch.softmax(logits, -1)
    q = torch.full((cfg.num_targets,), -np.log(cfg.num_targets), dtype=torch.float64) if prior is None else prior
    reg = w * (pr * (torch.log(pr) - q)).sum()                     # bilstm.py:258-260 / :265-267
    (ctc + reg).backward()
    assert abs(m.reg_loss.item() - reg.item()) < 2e-2 * abs(reg.item()) + 1e-3, (m.reg_loss.item(), reg.item())
    grads = {k: v.cpu().double() for k, v in m.to_tf_dict(grads=True).items()}
    bad = {k: ((v - p64[k].grad).norm() / (p64[k].grad.norm() + 1e-12)).item() for k, v in grads.items()}
    bad = {k: r for k, r in bad.items() if r > 5e-2}
    assert not bad, bad


def test_layer0_residual(cuda_dev):
    """input_dim == 2*num_projects switches on finput = finput + concat(fwd, bwd) in layer 0 (bilstm.py:199-200)."""
    from lstm_ctc_b200.model import AcousticModel
    cfg = oracle.OracleConfig(input_dim=64, num_layers=2, num_neurons=64, num_projects=32, num_targets=9, use_peepholes=True, num_experts=0)
    params = oracle.init_params(cfg, seed=51, bias_scale=0.1)
    x, lens, labels = make_batch(cfg, B=3, T=12, Lmax=4, seed=52)
    m = AcousticModel(nnet_config(cfg), cuda_dev, init=False)
    assert m.cfg.residual0
    m.from_tf_dict(params)
    loss_sum, _ = m.loss_and_grad(x.float().to(cuda_dev), lens.to(cuda_dev), labels.to(cuda_dev))
    p64 = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    ctc, _, ref_logits = oracle.training_loss(p64, cfg, x, lens, labels, 0.0)
    ctc.backward()
    logits = m._out_ws(12, 3)["logits"].cpu().double()
    live = (torch.arange(12).unsqueeze(0) < lens.unsqueeze(1))
    assert (logits - ref_logits.detach())[live].abs().max().item() < 1e-2 * ref_logits.abs().max().item()
    assert abs(loss_sum.item() - ctc.item()) < 2e-3 * abs(ctc.item())
    grads = {k: v.cpu().double() for k, v in m.to_tf_dict(grads=True).items()}
    bad = {k: ((v - p64[k].grad).norm() / (p64[k].grad.norm() + 1e-12)).item() for k, v in grads.items()}
    bad = {k: r for k, r in bad.items() if r > 5e-2}
    assert not bad, bad


@pytest.mark.parametrize("T", [1, 9])
def test_edge_lengths_through_the_whole_path(cuda_dev, T):
    """Ragged edge cases in ONE batch, through BiLSTM -> mixture -> CTC -> gradients: an empty utterance (sequence_length 0: zero
    output rows, skipped by CTC, loss 0 -- SURVEY 8c), a one-frame utterance, an utterance with no labels, one whose labels are
    longer than its frames (ignore_longer_outputs_than_inputs=True, graph.py:113: loss 0, gradient 0) and a full one; T = 1 is
    the shortest batch the kernels can be handed."""
    from lstm_ctc_b200 import _lib
    from lstm_ctc_b200.model import AcousticModel
    cfg = oracle.OracleConfig(**CASES["mos_k4"])
    cfg.num_layers = 2
    params = oracle.init_params(cfg, seed=61, bias_scale=0.1)
    B = 5
    g = torch.Generator().manual_seed(62)
    x = torch.randn(B, T, cfg.input_dim, generator=g, dtype=torch.float64)
    lens = torch.tensor([0, 1, T, max(1, T // 2), T], dtype=torch.int32)
    labels = torch.tensor([[1, -1, -1, -1, -1, -1], [2, -1, -1, -1, -1, -1], [-1, -1, -1, -1, -1, -1],
                           [3, 4, 5, 6, 7, 8], [1, 2, -1, -1, -1, -1]])
    if T == 1:
        labels[4, 1] = -1
    for b in range(B):
        x[b, lens[b]:] = 0
    p64 = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    ctc, _, ref_logits = oracle.training_loss(p64, cfg, x, lens, labels, l2_decay_weight=0.0)
    ctc.backward()
    m = AcousticModel(nnet_config(cfg), cuda_dev, init=False)
    m.from_tf_dict(params)
    loss_sum, loss = m.loss_and_grad(x.float().to(cuda_dev), lens.to(cuda_dev), labels.to(cuda_dev))
    torch.cuda.synchronize()
    assert _lib.lib().lcb_device_error(1) == 0
    loss = loss.cpu()
    assert loss[0].item() == 0.0                                   # empty utterance
    if T > 1:
        assert loss[3].item() == 0.0                               # 6 labels, T // 2 frames: ignored
    assert abs(loss_sum.item() - ctc.item()) < 2e-3 * max(abs(ctc.item()), 1.0), (loss_sum.item(), ctc.item())
    enc = m.enc._workspace(T, B, True)["Hout"][-1].float().view(T, B, -1)
    assert enc[:, 0].abs().max().item() == 0.0                     # sequence_length 0: every output row exactly zero
    assert enc[1:, 1].abs().max().item() == 0.0 if T > 1 else True
    grads = {k: v.cpu().double() for k, v in m.to_tf_dict(grads=True).items()}
    bad = {}
    for k, v in grads.items():
        rg = p64[k].grad
        if rg.norm().item() < 1e-9:
            assert v.norm().item() < 1e-6, k
            continue
        rel = ((v - rg).norm() / rg.norm()).item()
        if rel > 5e-2:
            bad[k] = rel
    assert not bad, bad


@pytest.mark.parametrize("B,T,keep", [(40, 120, 0.9), (64, 100, 1.0)])
def test_banded_mixture_backward_equals_serial(cuda_dev, B, T, keep):
    """AcousticModel.top_overlap: the mixture layer's backward runs in three bands of frames, outermost first, the inner two on a
    side stream, and the top layer's BPTT starts on the outer band as range launches [0,T/6), [T/6,T/3), [T/3,T).  d loss / d z of
    every layer is bit-identical to the serial order; the parameter gradients agree up to the fp32 summation order of the
    accumulating weight-gradient GEMMs."""
    from lstm_ctc_b200.model import AcousticModel
    cfg = oracle.OracleConfig(input_dim=24, num_layers=2, num_neurons=512, num_projects=512, num_targets=20, use_peepholes=True, num_experts=4)
    params = oracle.init_params(cfg, seed=31, bias_scale=0.1)
    x, lens, labels = make_batch(cfg, B=B, T=T, Lmax=10, seed=32)
    nc = nnet_config(cfg)
    nc["dropout_rate"] = keep
    res = []
    for top in (False, True):
        m = AcousticModel(nc, cuda_dev, init=False)
        m.from_tf_dict(params)
        m.top_overlap = top
        m.enc.debug_dz = {}
        loss_sum, _ = m.loss_and_grad(x.float().to(cuda_dev), lens.to(cuda_dev), labels.to(cuda_dev), seq_len_host=lens)
        torch.cuda.synchronize()
        res.append((float(loss_sum), m.params.gflat.clone(), {k: v.clone() for k, v in m.enc.debug_dz.items()}))
    from lstm_ctc_b200 import _lib
    assert _lib.lib().lcb_device_error(1) == 0
    (l0, g0, dz0), (l1, g1, dz1) = res
    assert l0 == l1
    assert sorted(dz0) == sorted(dz1) == [0, 1]
    for k in dz0:
        assert torch.equal(dz0[k], dz1[k])
    assert (g0 - g1).abs().max().item() <= 1e-5 * g0.abs().max().item()
    assert g0.abs().max().item() > 0
