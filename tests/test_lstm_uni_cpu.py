"""nnet_type 'lstm' (functional core of nnet/lstm.py:125-368), CPU side: the oracle's uni-directional residual stack, and the
embedding of its variables into the BiLSTM variable set with zero backward cells (lstm_ctc_b200/lstm.py)."""
import torch

import oracle
from lstm_ctc_b200.blstm import ModelConfig
from lstm_ctc_b200.lstm import (backward_half_is_zero, embed_uni_variables, extract_uni_variables, random_uni_variables,
                                uni_param_names)


def _cfgs(D=24, L=3, H=64, P=32, V=12, K=4):
    ocfg = oracle.OracleConfig(input_dim=D, num_layers=L, num_neurons=H, num_projects=P, num_targets=V, use_peepholes=True, num_experts=K)
    nc = {"nnet_type": "lstm", "input_dim": D, "num_layers": L, "num_neurons": H, "num_projects": P, "num_targets": V,
          "num_experts": K, "dropout_rate": 1.0}
    return ocfg, nc


def test_config_of_the_lstm_builder():
    _, nc = _cfgs()
    c = ModelConfig(nc)
    assert c.uni and c.use_peepholes and c.forget_bias == 1.0          # lstm.py:238-244: peepholes hard-coded, default forget bias
    assert [c.uni_residual(i) for i in range(3)] == [False, True, True]  # lstm.py:236: no residual on layer 0 when D != P
    c2 = ModelConfig(dict(nc, input_dim=32))
    assert [c2.uni_residual(i) for i in range(3)] == [True, True, True]
    cb = ModelConfig(dict(nc, nnet_type="blstm"))
    assert not cb.uni and cb.forget_bias == 5.0 and not any(cb.uni_residual(i) for i in range(3))


def test_embedding_round_trip_and_invariant():
    for K in (0, 4):
        ocfg, nc = _cfgs(K=K)
        c = ModelConfig(nc)
        p = oracle.init_lstm_params(ocfg, seed=3, bias_scale=0.1)
        assert list(p) == uni_param_names(c) == oracle.lstm_param_order(ocfg)
        bi = embed_uni_variables(c, p)
        assert set(bi) == set(oracle.param_order(ocfg))                  # exactly the BiLSTM variable set
        assert backward_half_is_zero(c, bi)
        back = extract_uni_variables(c, bi)
        for k in p:
            assert torch.equal(back[k].double(), p[k].float().double()), k
        bi["bd1/brnn1/bias"][0] = 1.0
        assert not backward_half_is_zero(c, bi)
        r = random_uni_variables(c, seed=5)
        assert list(r) == uni_param_names(c) and all(r[k].shape == p[k].shape for k in p)


def test_embedded_bilstm_emits_zero_backward_half():
    """the zero backward cells emit exactly 0 (so a BiLSTM over the embedded variables carries the uni-directional model)"""
    ocfg, nc = _cfgs(K=0)
    p = {k: v.float().double() for k, v in oracle.init_lstm_params(ocfg, seed=1, bias_scale=0.1).items()}   # (the embedding holds fp32)
    bi = {k: v.double() for k, v in embed_uni_variables(ModelConfig(nc), p).items()}
    x = torch.randn(3, 9, 24, dtype=torch.float64, generator=torch.Generator().manual_seed(7))
    lens = torch.tensor([9, 6, 2], dtype=torch.int32)
    ob = oracle.OracleConfig(input_dim=24, num_layers=1, num_neurons=64, num_projects=32, num_targets=12, use_peepholes=True,
                             forget_bias=1.0)
    enc, _ = oracle.blstm_forward(bi, ob, x, lens)                       # first layer only: no residual there (D != P)
    ou = oracle.OracleConfig(input_dim=24, num_layers=1, num_neurons=64, num_projects=32, num_targets=12, use_peepholes=True)
    uni = oracle.lstm_forward(p, ou, x, lens)
    assert enc[:, :, 32:].abs().max().item() == 0.0
    assert torch.allclose(enc[:, :, :32], uni, atol=1e-12)


def test_oracle_residual_masking_and_gradient():
    ocfg, _ = _cfgs(D=32, L=2, K=0)                                      # D == P: residual on layer 0 too
    p = {k: v.clone().requires_grad_(True) for k, v in oracle.init_lstm_params(ocfg, seed=2, bias_scale=0.1).items()}
    x = torch.randn(2, 6, 32, dtype=torch.float64, generator=torch.Generator().manual_seed(8))
    lens = torch.tensor([6, 3], dtype=torch.int32)
    out = oracle.lstm_forward(p, ocfg, x, lens)
    assert out[1, 3:].abs().max().item() == 0.0                           # dynamic_rnn zero output past sequence_length
    # residual: with a zero projection the cell adds nothing, so the stack is the identity on live frames
    pz = {k: (torch.zeros_like(v) if k.endswith("projection/kernel") else v.detach()) for k, v in p.items()}
    ident = oracle.lstm_forward(pz, ocfg, x, lens)
    assert torch.allclose(ident[0], x[0]) and torch.allclose(ident[1, :3], x[1, :3])
    labels = torch.tensor([[1, 2], [3, -1]])
    ctc, total, _ = oracle.lstm_training_loss(p, ocfg, x, lens, labels)
    total.backward()
    assert all(v.grad is not None and torch.isfinite(v.grad).all() for v in p.values())
