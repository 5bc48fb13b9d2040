"""`create_logits_lstm` -- the uni-directional residual LSTM stack of /root/reference/nnet/lstm.py:125-368 (nnet_type 'lstm').

The reference's builder does not run as shipped (it calls helpers that do not exist and an older create_moe); what
is built here is its functional core, with TF r1.8 semantics:
    cells[i] = DropoutWrapper([ResidualWrapper](LSTMCell(num_units, num_proj, use_peepholes=True)), output_keep_prob)
               -- no ResidualWrapper on layer 0 when input_dim != num_projects (lstm.py:236-260); forget_bias = 1.0 (default)
    tf.nn.dynamic_rnn(cells[i], ..., sequence_length, scope="drnn{i}")   one layer after the other (lstm.py:263-287)
    affine or mixture output layer over [N, num_projects] (lstm.py:318-345)
Feature projection, the ornn / orthogonality regularisers and batch-norm (helpers absent from the reference) are not part of it.

How it runs: on the SAME cluster-persistent kernels as the BiLSTM, as a BiLSTM whose backward cells -- and every weight that
reads their output -- are zero.  That state is exactly invariant under training: a zero backward cell emits h = m * 0 = 0, every
gradient that reaches it is multiplied by a zero weight, and the gradients of the zero weights are products with the zero
activations, so L2, clipping and SGD / Momentum / Adam all leave them at 0.0 (asserted by tests/test_lstm_uni_gpu.py).  The
recurrence kernels run in their native uni-directional mode (num_dirs = 1: only direction-0 clusters and column halves), so the
zero cells cost no recurrence time; the variables, checkpoints and gradients the caller sees are the uni-directional ones, under
the names TF gives them (drnn{i}/lstm_cell/{kernel,bias,w_*_diag,projection/kernel})."""
import math
from typing import Dict

import torch

from .blstm import ModelConfig


def uni_prefix(i):
    return "drnn%d/lstm_cell" % i


def uni_param_names(cfg: ModelConfig):
    names = []
    for i in range(cfg.num_layers):
        p = uni_prefix(i)
        names += [p + "/kernel", p + "/bias", p + "/w_f_diag", p + "/w_i_diag", p + "/w_o_diag", p + "/projection/kernel"]
    names += ["Variable", "Variable_1"] + (["Variable_2", "Variable_3"] if cfg.K > 0 else [])
    return names


def uni_din(cfg: ModelConfig, i):
    return cfg.input_dim if i == 0 else cfg.P


def embed_uni_variables(cfg: ModelConfig, uni: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """uni-directional variables -> the BiLSTM variable set (bilstm.py names) with zero backward cells."""
    H, P = cfg.H, cfg.P
    out = {}
    for i in range(cfg.num_layers):
        k = uni[uni_prefix(i) + "/kernel"].to(torch.float32)
        din_u, din_b = uni_din(cfg, i), cfg.din(i)
        assert k.shape == (din_u + P, 4 * H), (k.shape, din_u, P, H)
        kb = torch.zeros(din_b + P, 4 * H, dtype=torch.float32)
        kb[:din_u] = k[:din_u]                       # x part (layer > 0: the forward half of the concatenated input)
        kb[din_b:] = k[din_u:]                       # recurrent part
        f, b = "fd%d/frnn%d" % (i, i), "bd%d/brnn%d" % (i, i)
        out[f + "/kernel"] = kb
        out[b + "/kernel"] = torch.zeros_like(kb)
        for n, shape in (("bias", (4 * H,)), ("w_f_diag", (H,)), ("w_i_diag", (H,)), ("w_o_diag", (H,)), ("projection/kernel", (H, P))):
            v = uni[uni_prefix(i) + "/" + n].to(torch.float32)
            assert tuple(v.shape) == shape, (n, v.shape, shape)
            out[f + "/" + n] = v.clone()
            out[b + "/" + n] = torch.zeros(shape, dtype=torch.float32)
    for n in (("Variable", "Variable_2") if cfg.K > 0 else ("Variable",)):       # [P, .] -> [2P, .] with zero rows for the backward half
        w = uni[n].to(torch.float32)
        assert w.shape[0] == P, (n, w.shape)
        out[n] = torch.cat([w, torch.zeros_like(w)], 0)
    for n in (("Variable_1", "Variable_3") if cfg.K > 0 else ("Variable_1",)):
        out[n] = uni[n].to(torch.float32).clone()
    return out


def extract_uni_variables(cfg: ModelConfig, bi: Dict[str, torch.Tensor]) -> Dict[str, torch.Tensor]:
    """inverse of embed_uni_variables (also for gradients)"""
    P = cfg.P
    out = {}
    for i in range(cfg.num_layers):
        f = "fd%d/frnn%d" % (i, i)
        din_u, din_b = uni_din(cfg, i), cfg.din(i)
        kb = bi[f + "/kernel"]
        out[uni_prefix(i) + "/kernel"] = torch.cat([kb[:din_u], kb[din_b:]], 0).clone()
        for n in ("bias", "w_f_diag", "w_i_diag", "w_o_diag", "projection/kernel"):
            out[uni_prefix(i) + "/" + n] = bi[f + "/" + n].clone()
    for n in (("Variable", "Variable_2") if cfg.K > 0 else ("Variable",)):
        out[n] = bi[n][:P].clone()
    for n in (("Variable_1", "Variable_3") if cfg.K > 0 else ("Variable_1",)):
        out[n] = bi[n].clone()
    return out


def backward_half_is_zero(cfg: ModelConfig, bi: Dict[str, torch.Tensor]) -> bool:
    """True iff every backward-cell variable and every weight reading the backward half is exactly 0 (the invariant)."""
    P = cfg.P
    for i in range(cfg.num_layers):
        b = "bd%d/brnn%d" % (i, i)
        for k, v in bi.items():
            if k.startswith(b + "/") and bool((v != 0).any()):
                return False
        if i > 0 and bool((bi["fd%d/frnn%d/kernel" % (i, i)][P:2 * P] != 0).any()):
            return False
    for n in (("Variable", "Variable_2") if cfg.K > 0 else ("Variable",)):
        if bool((bi[n][P:] != 0).any()):
            return False
    return True


def random_uni_variables(cfg: ModelConfig, seed=None) -> Dict[str, torch.Tensor]:
    """TF default initialisers: glorot-uniform kernels / peepholes / projection, zero biases; truncated-normal output layer
    with stddev 1/sqrt(output_dim) (lstm.py:331-337)."""
    from .model import _glorot, _trunc_normal
    g = torch.Generator()
    if seed is None:
        g.seed()
    else:
        g.manual_seed(int(seed))
    H, P = cfg.H, cfg.P
    tf = {}
    for i in range(cfg.num_layers):
        p = uni_prefix(i)
        tf[p + "/kernel"] = _glorot((uni_din(cfg, i) + P, 4 * H), g)
        tf[p + "/bias"] = torch.zeros(4 * H)
        for w in ("w_f_diag", "w_i_diag", "w_o_diag"):
            tf[p + "/" + w] = _glorot((H,), g)
        tf[p + "/projection/kernel"] = _glorot((H, P), g)
    std = 1.0 / math.sqrt(P)
    if cfg.K > 0:
        tf["Variable"] = _trunc_normal((P, cfg.K), std, g)
        tf["Variable_1"] = torch.zeros(cfg.K)
        tf["Variable_2"] = _trunc_normal((P, cfg.K * cfg.V), std, g)
        tf["Variable_3"] = torch.zeros(cfg.K * cfg.V)
    else:
        tf["Variable"] = _trunc_normal((P, cfg.V), std, g)
        tf["Variable_1"] = torch.zeros(cfg.V)
    return tf


def create_logits_lstm(nnet_input, sequence_length, nnet_config, model=None):
    """Reference signature and return pair (lstm.py:125,368): (logits [B,T,V], reg_loss).  reg_loss is None: the regularisers of
    the reference builder (ornn, orthogonality, feature projection) have no definition in its tree."""
    from .graph import _default_model
    from .model import AcousticModel
    cfg = dict(nnet_config)
    cfg["nnet_type"] = "lstm"
    m = model or _default_model[0]
    if m is None:
        m = AcousticModel(cfg)
        _default_model[0] = m
    training = cfg.get("is_training")
    training = True if training is None else bool(training)
    return m.forward_logits(nnet_input, sequence_length, training=training), None
