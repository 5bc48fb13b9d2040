"""Pins for the LSTM / optimizer part of the oracle (oracle/model.py), which the reference's tree cannot pin (no tests, no TF):
an INDEPENDENT implementation of the same published semantics -- torch.nn.LSTM (cuDNN-style fused cell, gate order i,f,g,o,
proj_size, packed ragged sequences, bidirectional stacking), torch.optim.Adam / SGD, torch.nn.utils.clip_grad_norm_ -- must agree
with the restatement of TF r1.8's LSTMCell / dynamic_rnn / reverse_sequence / AdamOptimizer / clip_by_global_norm after the
documented re-mapping (TF gate order i,j,f,o; forget_bias folded into the bias; kernel = [W_x ; W_h]^T).  Peepholes have no
torch counterpart: they are covered by finite differences (gradcheck) and by the closed form of one step."""
import math

import torch
from torch.nn.utils.rnn import pack_padded_sequence, pad_packed_sequence

import oracle
from oracle import model as om


def _to_torch_lstm(params, cfg, bidirectional=True):
    """Load oracle variables (no peepholes) into torch.nn.LSTM(num_layers, bidirectional, proj_size)."""
    H, P = cfg.num_neurons, cfg.num_projects
    net = torch.nn.LSTM(cfg.input_dim, H, num_layers=cfg.num_layers, batch_first=True, bidirectional=bidirectional, proj_size=P).double()
    order = [0, 2, 1, 3]                                   # TF blocks (i, j, f, o) -> torch rows (i, f, g, o)
    with torch.no_grad():
        for i in range(cfg.num_layers):
            for d, (dn, cn) in enumerate((("fd", "frnn"), ("bd", "brnn"))[: 2 if bidirectional else 1]):
                pre = "%s%d/%s%d" % (dn, i, cn, i)
                k = params[pre + "/kernel"]                # [Din + P, 4H]
                din = k.shape[0] - P
                blocks = torch.chunk(k, 4, dim=1)
                w = torch.cat([blocks[j] for j in order], 1).t()        # [4H, Din + P]
                b = torch.chunk(params[pre + "/bias"], 4)
                bias = torch.cat([b[0], b[2] + cfg.forget_bias, b[1], b[3]])
                sfx = "_l%d%s" % (i, "_reverse" if d == 1 else "")
                getattr(net, "weight_ih" + sfx).copy_(w[:, :din])
                getattr(net, "weight_hh" + sfx).copy_(w[:, din:])
                getattr(net, "bias_ih" + sfx).copy_(bias)
                getattr(net, "bias_hh" + sfx).zero_()
                getattr(net, "weight_hr" + sfx).copy_(params[pre + "/projection/kernel"].t())
    return net


def test_bilstm_stack_agrees_with_torch_nn_lstm():
    """3-layer BiLSTM with projection, ragged lengths: outputs (zero past sequence_length), final states and input gradient."""
    cfg = oracle.OracleConfig(input_dim=10, num_layers=3, num_neurons=12, num_projects=6, num_targets=5, use_peepholes=False)
    params = oracle.init_params(cfg, seed=4, bias_scale=0.3)
    g = torch.Generator().manual_seed(5)
    B, T = 4, 9
    x = torch.randn(B, T, cfg.input_dim, generator=g, dtype=torch.float64, requires_grad=True)
    lens = torch.tensor([9, 7, 4, 1], dtype=torch.int32)
    out, enc = oracle.blstm_forward(params, cfg, x, lens)
    net = _to_torch_lstm(params, cfg)
    x2 = x.detach().clone().requires_grad_(True)
    packed = pack_padded_sequence(x2, lens.to(torch.int64), batch_first=True, enforce_sorted=True)
    yp, (hn, cn) = net(packed)
    y, _ = pad_packed_sequence(yp, batch_first=True, total_length=T)
    assert torch.allclose(out, y, atol=1e-12)
    for b in range(B):
        assert out[b, lens[b]:].abs().max().item() == 0.0 if lens[b] < T else True
    # encoder = concat(c_fw, h_fw, c_bw, h_bw) of the LAST layer (bilstm.py:206-208)
    H, P = cfg.num_neurons, cfg.num_projects
    ref_enc = torch.cat([cn[-2], hn[-2], cn[-1], hn[-1]], 1)
    assert torch.allclose(enc, ref_enc, atol=1e-12)
    w = torch.randn(out.shape, generator=g, dtype=torch.float64)
    (out * w).sum().backward()
    (y * w).sum().backward()
    assert torch.allclose(x.grad, x2.grad, atol=1e-11)


def test_uni_stack_without_residual_agrees_with_torch_nn_lstm():
    """dynamic_rnn + LSTMCell alone (one direction, forget_bias 1.0): the building block of oracle.lstm_forward."""
    cfg = oracle.OracleConfig(input_dim=7, num_layers=1, num_neurons=9, num_projects=5, num_targets=4, use_peepholes=False, forget_bias=1.0)
    params = oracle.init_params(cfg, seed=6, bias_scale=0.2)
    g = torch.Generator().manual_seed(7)
    x = torch.randn(3, 8, 7, generator=g, dtype=torch.float64)
    lens = torch.tensor([8, 5, 2], dtype=torch.int32)
    cellp = om._cell_params(params, cfg, 0, "fd", "frnn")
    out, (c, h) = om.dynamic_rnn(x, lens, cellp, cfg.forget_bias)
    net = _to_torch_lstm(params, cfg, bidirectional=False)
    yp, (hn, cn) = net(pack_padded_sequence(x, lens.to(torch.int64), batch_first=True))
    y, _ = pad_packed_sequence(yp, batch_first=True, total_length=8)
    assert torch.allclose(out, y, atol=1e-12) and torch.allclose(h, hn[0], atol=1e-12) and torch.allclose(c, cn[0], atol=1e-12)
    # the residual variant adds the input inside the wrapper: out = x + h on live frames, 0 past the length
    cfg2 = oracle.OracleConfig(input_dim=5, num_layers=1, num_neurons=9, num_projects=5, num_targets=4, use_peepholes=False, forget_bias=1.0)
    p2 = oracle.init_params(cfg2, seed=8)
    x5 = torch.randn(3, 8, 5, generator=g, dtype=torch.float64)
    cp2 = om._cell_params(p2, cfg2, 0, "fd", "frnn")
    plain, _ = om.dynamic_rnn(x5, lens, cp2, 1.0)
    res, _ = om.dynamic_rnn(x5, lens, cp2, 1.0, residual=True)
    live = (torch.arange(8).unsqueeze(0) < lens.unsqueeze(1)).unsqueeze(2).double()
    assert torch.allclose(res, plain + x5 * live, atol=1e-14)


def test_peephole_cell_closed_form_and_gradcheck():
    """One step of LSTMCell.call with peepholes against the formulas of SURVEY 8a3 written out independently, and finite
    differences through two steps."""
    g = torch.Generator().manual_seed(9)
    B, D, H, P = 2, 3, 4, 3
    r = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64) * 0.5
    x, c0, h0 = r(B, D), r(B, H), r(B, P)
    kernel, bias, wf, wi, wo, proj = r(D + P, 4 * H), r(4 * H), r(H), r(H), r(H), r(H, P)
    c1, h1 = om.lstm_cell(x, c0, h0, kernel, bias, wf, wi, wo, proj, 5.0)
    z = torch.cat([x, h0], 1) @ kernel + bias
    zi, zj, zf, zo = z[:, :H], z[:, H:2 * H], z[:, 2 * H:3 * H], z[:, 3 * H:]
    sig = lambda v: 1 / (1 + torch.exp(-v))
    c_ref = sig(zf + 5.0 + wf * c0) * c0 + sig(zi + wi * c0) * torch.tanh(zj)
    m_ref = sig(zo + wo * c_ref) * torch.tanh(c_ref)
    assert torch.allclose(c1, c_ref, atol=1e-14) and torch.allclose(h1, m_ref @ proj, atol=1e-14)

    def two_steps(x, kernel, bias, wf, wi, wo, proj):
        c, h = om.lstm_cell(x, c0, h0, kernel, bias, wf, wi, wo, proj, 5.0)
        c, h = om.lstm_cell(x * 0.5, c, h, kernel, bias, wf, wi, wo, proj, 5.0)
        return h.sum() + (c * c).sum()
    args = [t.clone().requires_grad_(True) for t in (x, kernel, bias, wf, wi, wo, proj)]
    assert torch.autograd.gradcheck(two_steps, args, eps=1e-6, atol=1e-6)


def test_reverse_sequence_matches_definition():
    x = torch.arange(2 * 5 * 1, dtype=torch.float64).view(2, 5, 1)
    y = om.reverse_sequence(x, torch.tensor([3, 5], dtype=torch.int32))
    assert y[0, :, 0].tolist() == [2.0, 1.0, 0.0, 3.0, 4.0]          # reversed inside the length, untouched past it
    assert y[1, :, 0].tolist() == [9.0, 8.0, 7.0, 6.0, 5.0]


def test_clip_and_optimizers_agree_with_torch():
    g = torch.Generator().manual_seed(10)
    p = {"a/kernel": torch.randn(5, 4, generator=g, dtype=torch.float64), "a/bias": torch.randn(4, generator=g, dtype=torch.float64)}
    grads = {k: torch.randn(v.shape, generator=g, dtype=torch.float64) * 3 for k, v in p.items()}
    clipped, gn = oracle.clip_by_global_norm(grads, 5.0)
    tp = [torch.nn.Parameter(v.clone()) for v in p.values()]
    for q, gr in zip(tp, grads.values()):
        q.grad = gr.clone()
    tn = torch.nn.utils.clip_grad_norm_(tp, 5.0)
    assert abs(float(tn) - float(gn)) < 1e-12
    for q, k in zip(tp, p):
        assert torch.allclose(q.grad, clipped[k], rtol=1e-6)          # (torch divides by norm + 1e-6)
    # TF Adam: lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t), update = lr_t * m / (sqrt(v) + eps); torch puts eps on sqrt(v_hat):
    # the two differ only through eps, i.e. by O(eps / sqrt(v)) -- compare with eps where that is negligible
    opt = torch.optim.Adam(tp, lr=1e-2, betas=(0.9, 0.999), eps=1e-12)
    state, cur = {}, {k: v.clone() for k, v in p.items()}
    for step in range(4):
        gs = {k: torch.randn(v.shape, generator=g, dtype=torch.float64) for k, v in p.items()}
        for q, k in zip(tp, p):
            q.grad = gs[k].clone()
        opt.step()
        cur = oracle.adam_step(cur, gs, state, 1e-2, eps=1e-12)
        for q, k in zip(tp, p):
            assert torch.allclose(q.detach(), cur[k], atol=1e-9), (step, k)
    # Momentum: accum = momentum * accum + g; w -= lr * accum  (tf.train.MomentumOptimizer) == torch SGD(momentum, dampening 0)
    tp2 = [torch.nn.Parameter(v.clone()) for v in p.values()]
    sgd = torch.optim.SGD(tp2, lr=1e-2, momentum=0.9)
    state, cur = {}, {k: v.clone() for k, v in p.items()}
    for step in range(3):
        gs = {k: torch.randn(v.shape, generator=g, dtype=torch.float64) for k, v in p.items()}
        for q, k in zip(tp2, p):
            q.grad = gs[k].clone()
        sgd.step()
        cur = oracle.momentum_step(cur, gs, state, 1e-2, momentum=0.9)
        for q, k in zip(tp2, p):
            assert torch.allclose(q.detach(), cur[k], atol=1e-12), (step, k)


def test_l2_skips_only_names_containing_bias():
    p = {"fd0/frnn0/kernel": torch.ones(2, 2), "fd0/frnn0/bias": torch.ones(3) * 7, "Variable_1": torch.ones(4) * 2}
    # graph.py:183-189: LSTM '.../bias' is skipped, the unnamed output-layer bias Variable_1 is decayed
    assert abs(float(oracle.l2_loss(p, 0.5)) - 0.5 * (4 * 1 + 4 * 4) / 2) < 1e-12


def test_greedy_decode_and_edit_distance_known_answers():
    V = 4                                                   # blank = 3
    seq = [0, 0, 3, 0, 1, 1, 3, 3, 2]                       # collapse repeats, drop blanks -> 0 0 1 2
    logits = torch.full((1, len(seq), V), -5.0)
    for t, s in enumerate(seq):
        logits[0, t, s] = 5.0
    hyp = oracle.greedy_decode(logits, torch.tensor([len(seq)], dtype=torch.int32))
    assert list(hyp[0]) == [0, 0, 1, 2]
    assert list(oracle.greedy_decode(logits, torch.tensor([4], dtype=torch.int32))[0]) == [0, 0]
    assert oracle.edit_distance([0, 0, 1, 2], [0, 1, 2]) == 1 and oracle.edit_distance([], [1, 2]) == 2
    assert oracle.edit_distance([1, 2, 3], [1, 2, 3]) == 0 and oracle.edit_distance([1, 2, 3], [3, 2, 1]) == 2


def test_mixture_output_layer_against_explicit_loops():
    """create_moe (moe.py:29-72): y[n,v] = sum_k softmax_k(x Wp + bp)[n,k] * tau * tanh(x W + b)[n, k*V + v]  -- the expert
    logits are k-major columns (moe.py:60), the mixture is over tanh-bounded LOGITS; dropout acts on the mixture
    weights and on the expert logits with keep-probability semantics (moe.py:46,61)."""
    g = torch.Generator().manual_seed(11)
    N, D, K, V, tau = 3, 5, 4, 6, 10.0
    r = lambda *s: torch.randn(*s, generator=g, dtype=torch.float64)
    x, Wp, bp, W, b = r(N, D), r(D, K), r(K), r(D, K * V), r(K * V)
    y = oracle.create_moe(x, Wp, bp, W, b, V, K, tau)
    ref = torch.zeros(N, V, dtype=torch.float64)
    for n in range(N):
        a = [sum(x[n, d] * Wp[d, k] for d in range(D)) + bp[k] for k in range(K)]
        mx = max(a)
        e = [math.exp(float(v - mx)) for v in a]
        pi = [v / sum(e) for v in e]
        for v in range(V):
            for k in range(K):
                z = sum(x[n, d] * W[d, k * V + v] for d in range(D)) + b[k * V + v]
                ref[n, v] += pi[k] * tau * math.tanh(float(z))
    assert torch.allclose(y, ref, atol=1e-12)
    assert y.abs().max().item() <= tau + 1e-9                         # |y| <= tau * sum_k pi_k = tau
    keep = 0.5
    mp = (torch.rand(N, K, 1, generator=g) < keep).double()
    md = (torch.rand(N, K, V, generator=g) < keep).double()
    yd = oracle.create_moe(x, Wp, bp, W, b, V, K, tau, keep_prob=keep, mask_prior=mp, mask_dec=md)
    pi = torch.softmax(x @ Wp + bp, 1).unsqueeze(2) * mp / keep
    dec = (tau * torch.tanh(x @ W + b)).reshape(N, K, V) * md / keep
    assert torch.allclose(yd, (pi * dec).sum(1), atol=1e-12)


def test_twin_without_rounding_is_the_oracle(monkeypatch):
    """oracle/twin.py restates the stack in the device's formulation (hoisted x-projection, folded recurrent weight W' =
    W_proj W_h acting on m_{t-1}, bulk projection after the loop).  With its rounding hooks switched off it must BE the oracle:
    outputs, gradients and per-step dz."""
    from oracle import twin
    monkeypatch.setattr(twin, "q16", lambda x: x)
    monkeypatch.setattr(twin, "gq_bf16", lambda x: x)
    cfg = oracle.OracleConfig(input_dim=12, num_layers=3, num_neurons=24, num_projects=16, num_targets=9, use_peepholes=True,
                              num_experts=3)
    p = oracle.init_params(cfg, seed=5, bias_scale=0.2)
    g = torch.Generator().manual_seed(6)
    B, T = 5, 11
    x = torch.randn(B, T, 12, generator=g, dtype=torch.float64)
    lens = torch.tensor([11, 7, 3, 11, 1], dtype=torch.int32)
    for b in range(B):
        x[b, lens[b]:] = 0
    keep = 0.8
    masks = {(i, d): (torch.rand(B, T, 16, generator=g) < keep).double() for i in range(3) for d in "fb"}
    mp = (torch.rand(B * T, 3, 1, generator=g) < keep).double()
    md = (torch.rand(B * T, 3, 9, generator=g) < keep).double()
    res = []
    for which in (0, 1):
        pr = {k: v.clone().requires_grad_(True) for k, v in p.items()}
        tr = {}
        if which == 0:
            enc, _ = oracle.blstm_forward(pr, cfg, x, lens, keep_prob=keep, masks=masks, trace=tr)
            y = oracle.create_moe(enc.reshape(-1, 32), pr["Variable"], pr["Variable_1"], pr["Variable_2"], pr["Variable_3"], 9, 3,
                                  cfg.moe_temp, keep_prob=keep, mask_prior=mp, mask_dec=md).reshape(B, T, 9)
        else:
            enc = twin.blstm_forward_twin(pr, cfg, x, lens, keep_prob=keep, masks=masks, trace=tr)
            y = twin.output_layer_twin(pr, cfg, enc, keep_prob=keep, mask_prior=mp, mask_dec=md)
        (y * torch.arange(y.numel(), dtype=torch.float64).reshape(y.shape).cos()).sum().backward()
        dz = torch.stack([z.grad for z in tr[(2, "b")]], 1)
        res.append((y.detach(), {k: v.grad for k, v in pr.items()}, dz))
    (y0, g0, dz0), (y1, g1, dz1) = res
    assert (y0 - y1).abs().max().item() < 1e-11
    assert (dz0 - dz1).abs().max().item() < 1e-11
    for k in g0:
        assert (g0[k] - g1[k]).abs().max().item() < 1e-10 * (1 + g0[k].abs().max().item()), k


def test_sensitivity_to_fp16_rounding():
    """Why long-sequence parity is checked against the precision twin: rounding ONLY the weights and the input features to fp16,
    once, moves the fp64 oracle's own outputs and gradients by an amount that grows steeply with T (a random peephole BiLSTM with
    forget bias 5 has exploding gradients: |g| grows ~10x per doubling of T).  Recorded in profiles/r02_oracle_fp16_sensitivity.txt
    at C1 dimensions; here a smaller stack, asserting the growth."""
    cfg = oracle.OracleConfig(input_dim=40, num_layers=3, num_neurons=128, num_projects=128, num_targets=20, use_peepholes=True)
    p = oracle.init_params(cfg, seed=101, bias_scale=0.1)
    rel = {}
    for T in (24, 96):
        g = torch.Generator().manual_seed(202)
        x = torch.randn(4, T, 40, generator=g, dtype=torch.float64)
        lens = torch.full((4,), T, dtype=torch.int32)
        lab = torch.randint(0, 19, (4, T // 8), generator=g)
        grads = []
        for half in (False, True):
            pr = {k: (v.half().double() if half else v.clone()).requires_grad_(True) for k, v in p.items()}
            ctc, _, _ = oracle.training_loss(pr, cfg, x.half().double() if half else x, lens, lab, l2_decay_weight=0.0)
            ctc.backward()
            grads.append({k: v.grad for k, v in pr.items()})
        rel[T] = max(((grads[1][k] - grads[0][k]).norm() / grads[0][k].norm()).item() for k in grads[0])
    assert rel[24] < 2e-2 and rel[96] > 3 * rel[24], rel
