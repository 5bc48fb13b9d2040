"""oracle/model.py -- TEST INFRASTRUCTURE.  PyTorch-CPU restatement (fp64 for parity, fp32 for
the timed CPU baseline) of the reference's graph for the hot path.  Deliberately written the
way the reference executes it (per-time-step cell, reverse_sequence copies, materialised
[N,K,V] expert tensor) -- it is the checker and the "reference CPU path", never the product.

Follows, line by line:
  /root/reference/nnet/bilstm.py:104-273   create_logits_blstm
  /root/reference/nnet/moe.py:29-72        create_moe
  /root/reference/nnet/graph.py:51-209     loss / training graph
  /root/reference/nnet/lstm.py:125-368     create_logits_lstm -- its functional core only (lstm_* functions at the end)
and the TF r1.8 semantics those call into (third-party, restated from the published sources):
  rnn_cell_impl.LSTMCell.call  (gate order i,j,f,o; peepholes; forget_bias; num_proj)
  rnn.dynamic_rnn / _rnn_step  (zero output + state copy-through past sequence_length)
  array_ops.reverse_sequence, clip_ops.clip_by_global_norm, training/adam.py
"""
import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional

import numpy as np
import torch

from .ctc import ctc_loss_grad


@dataclass
class OracleConfig:
    input_dim: int
    num_layers: int
    num_neurons: int
    num_projects: int
    num_targets: int
    use_peepholes: bool = False
    num_experts: int = 0
    moe_temp: float = 10.0
    forget_bias: float = 5.0          # bilstm.py:133,154

    @staticmethod
    def from_nnet_config(c: dict) -> "OracleConfig":
        ctx = 1 + (c.get("left_context") or 0) + (c.get("right_context") or 0)
        return OracleConfig(
            input_dim=c["input_dim"] * ctx, num_layers=c["num_layers"],
            num_neurons=c["num_neurons"], num_projects=c["num_projects"],
            num_targets=c["num_targets"], use_peepholes=bool(c.get("use_peepholes") or False),
            num_experts=int(c.get("num_experts") or 0),
            moe_temp=float(c.get("moe_temp") if c.get("moe_temp") is not None else 10.0))


def param_order(cfg: OracleConfig) -> List[str]:
    """Variable names in TF creation order (trainable_variables order)."""
    names = []
    for i in range(cfg.num_layers):
        for d, c in (("fd", "frnn"), ("bd", "brnn")):
            p = "%s%d/%s%d" % (d, i, c, i)
            names += [p + "/kernel", p + "/bias"]
            if cfg.use_peepholes:
                names += [p + "/w_f_diag", p + "/w_i_diag", p + "/w_o_diag"]
            names += [p + "/projection/kernel"]
    if cfg.num_experts > 0:
        names += ["Variable", "Variable_1", "Variable_2", "Variable_3"]   # Wp, bp, W, b (moe.py:34-58)
    else:
        names += ["Variable", "Variable_1"]                                # W, b (bilstm.py:240-248)
    return names


def _glorot(shape, gen, dtype):
    fan_in, fan_out = (shape[0], shape[1]) if len(shape) == 2 else (shape[0], shape[0])
    lim = math.sqrt(6.0 / (fan_in + fan_out))
    return (torch.rand(shape, generator=gen, dtype=torch.float64) * 2 - 1).mul_(lim).to(dtype)


def _trunc_normal(shape, std, gen, dtype):
    x = torch.randn(shape, generator=gen, dtype=torch.float64)
    for _ in range(8):
        bad = x.abs() > 2
        if not bad.any():
            break
        x[bad] = torch.randn(int(bad.sum()), generator=gen, dtype=torch.float64)
    return x.clamp_(-2, 2).mul_(std).to(dtype)


def init_params(cfg: OracleConfig, seed=0, dtype=torch.float64, bias_scale=0.0) -> Dict[str, torch.Tensor]:
    """glorot-uniform LSTM kernels/peepholes/projection, zero biases, trunc-normal output layer
    (TF defaults; nnet-init.py:73).  bias_scale>0 randomises biases so tests exercise them."""
    g = torch.Generator().manual_seed(seed)
    H, P = cfg.num_neurons, cfg.num_projects
    p: Dict[str, torch.Tensor] = {}
    for i in range(cfg.num_layers):
        din = cfg.input_dim if i == 0 else 2 * P
        for d, c in (("fd", "frnn"), ("bd", "brnn")):
            pre = "%s%d/%s%d" % (d, i, c, i)
            p[pre + "/kernel"] = _glorot((din + P, 4 * H), g, dtype)
            p[pre + "/bias"] = (torch.randn(4 * H, generator=g, dtype=torch.float64) * bias_scale).to(dtype)
            if cfg.use_peepholes:
                for w in ("w_f_diag", "w_i_diag", "w_o_diag"):
                    p[pre + "/" + w] = _glorot((H,), g, dtype)
            p[pre + "/projection/kernel"] = _glorot((H, P), g, dtype)
    od = 2 * P
    if cfg.num_experts > 0:
        K, V = cfg.num_experts, cfg.num_targets
        std = 1.0 / math.sqrt(od)
        p["Variable"] = _trunc_normal((od, K), std, g, dtype)
        p["Variable_1"] = (torch.randn(K, generator=g, dtype=torch.float64) * bias_scale).to(dtype)
        p["Variable_2"] = _trunc_normal((od, K * V), std, g, dtype)
        p["Variable_3"] = (torch.randn(K * V, generator=g, dtype=torch.float64) * bias_scale).to(dtype)
    else:
        std = 1.0 / math.sqrt(H)
        p["Variable"] = _trunc_normal((od, cfg.num_targets), std, g, dtype)
        p["Variable_1"] = (torch.randn(cfg.num_targets, generator=g, dtype=torch.float64) * bias_scale).to(dtype)
    assert list(p.keys()) == param_order(cfg)
    return p


def reverse_sequence(x, seq_len):
    """tf.reverse_sequence(x, len, seq_axis=1, batch_axis=0) (bilstm.py:112,190,203)."""
    B, T = x.shape[0], x.shape[1]
    t = torch.arange(T).unsqueeze(0).expand(B, T)
    L = seq_len.to(torch.long).unsqueeze(1)
    idx = torch.where(t < L, L - 1 - t, t)
    return torch.gather(x, 1, idx.unsqueeze(-1).expand_as(x))


def lstm_cell(x_t, c, h, kernel, bias, w_f, w_i, w_o, proj, forget_bias, zs=None):
    """TF r1.8 rnn_cell_impl.LSTMCell.call.  zs: optional list that receives the pre-activation z_t (with its gradient
    retained), so tests can compare d loss / d z_t -- what the BPTT kernel emits -- layer by layer."""
    z = torch.cat([x_t, h], 1) @ kernel + bias
    if zs is not None:
        if z.requires_grad:
            z.retain_grad()
        zs.append(z)
    i, j, f, o = torch.chunk(z, 4, dim=1)
    if w_f is not None:
        c_new = torch.sigmoid(f + forget_bias + w_f * c) * c + torch.sigmoid(i + w_i * c) * torch.tanh(j)
        m = torch.sigmoid(o + w_o * c_new) * torch.tanh(c_new)
    else:
        c_new = torch.sigmoid(f + forget_bias) * c + torch.sigmoid(i) * torch.tanh(j)
        m = torch.sigmoid(o) * torch.tanh(c_new)
    h_new = m @ proj if proj is not None else m
    return c_new, h_new


def dynamic_rnn(x, seq_len, cellp, forget_bias, keep_prob=1.0, masks=None, residual=False, zs=None):
    """tf.nn.dynamic_rnn over a DropoutWrapper(LSTMCell) (bilstm.py:127-137,171-188):
    zero output and state copy-through for t >= seq_len[b]; dropout on the emitted output only.
    masks: optional [B,T,P] 0/1 tensor standing in for TF's (unmatchable) RNG stream."""
    B, T, _ = x.shape
    kernel, bias, w_f, w_i, w_o, proj = cellp
    H = kernel.shape[1] // 4
    P = proj.shape[1] if proj is not None else H
    c = x.new_zeros(B, H)
    h = x.new_zeros(B, P)
    outs = []
    for t in range(T):
        c_new, h_new = lstm_cell(x[:, t], c, h, kernel, bias, w_f, w_i, w_o, proj, forget_bias, zs)
        live = (t < seq_len).to(x.dtype).unsqueeze(1)
        out = h_new + x[:, t] if residual else h_new        # ResidualWrapper sits INSIDE the DropoutWrapper (lstm.py:247-259)
        if keep_prob < 1.0 and masks is not None:
            out = out * masks[:, t] / keep_prob
        outs.append(out * live)
        c = live * c_new + (1 - live) * c
        h = live * h_new + (1 - live) * h
    return torch.stack(outs, 1), (c, h)


def _cell_params(p, cfg, i, d, c):
    pre = "%s%d/%s%d" % (d, i, c, i)
    pe = cfg.use_peepholes
    return (p[pre + "/kernel"], p[pre + "/bias"],
            p[pre + "/w_f_diag"] if pe else None, p[pre + "/w_i_diag"] if pe else None,
            p[pre + "/w_o_diag"] if pe else None, p[pre + "/projection/kernel"])


def blstm_forward(p, cfg: OracleConfig, nnet_input, seq_len, keep_prob=1.0, masks=None, trace=None):
    """create_logits_blstm up to the encoder output (bilstm.py:104-211).
    Returns (output [B,T,2P], encoder [B,2(H+P)]).
    masks: optional dict {(layer, 'f'|'b'): [B,T,P]} in each direction's OWN time order.
    trace: optional dict that receives {(layer, 'f'|'b'): [z_0 .. z_{T-1}]}, the cells' pre-activations in each direction's own
    time order (see lstm_cell)."""
    finput = nnet_input
    binput = reverse_sequence(nnet_input, seq_len)
    fw_state = bw_state = None
    for i in range(cfg.num_layers):
        fo, fw_state = dynamic_rnn(finput, seq_len, _cell_params(p, cfg, i, "fd", "frnn"),
                                   cfg.forget_bias, keep_prob, None if masks is None else masks.get((i, "f")),
                                   zs=None if trace is None else trace.setdefault((i, "f"), []))
        bo, bw_state = dynamic_rnn(binput, seq_len, _cell_params(p, cfg, i, "bd", "brnn"),
                                   cfg.forget_bias, keep_prob, None if masks is None else masks.get((i, "b")),
                                   zs=None if trace is None else trace.setdefault((i, "b"), []))
        rbo = reverse_sequence(bo, seq_len)
        cat = torch.cat([fo, rbo], 2)
        if i == 0 and cfg.input_dim == 2 * cfg.num_projects:      # bilstm.py:199-200
            finput = finput + cat
        else:
            finput = cat
        binput = reverse_sequence(finput, seq_len)
    encoder = torch.cat([torch.cat(fw_state, 1), torch.cat(bw_state, 1)], 1)   # bilstm.py:206-208
    return finput, encoder


def create_moe(x, Wp, bp, W, b, num_targets, num_experts, tau, keep_prob=1.0, mask_prior=None, mask_dec=None):
    """nnet/moe.py:29-72 (materialises the [N,K,V] tensor exactly like the reference)."""
    y_prior = torch.softmax(x @ Wp + bp, dim=1).unsqueeze(2)                   # moe.py:43-45
    if keep_prob < 1.0 and mask_prior is not None:
        y_prior = y_prior * mask_prior / keep_prob                             # moe.py:46
    y_dec = tau * torch.tanh(x @ W + b)                                        # moe.py:59
    y_dec = y_dec.reshape(-1, num_experts, num_targets)                        # moe.py:60
    if keep_prob < 1.0 and mask_dec is not None:
        y_dec = y_dec * mask_dec / keep_prob                                   # moe.py:61
    return (y_prior * y_dec).sum(1)                                            # moe.py:71


def output_layer(p, cfg: OracleConfig, enc_out):
    """bilstm.py:227-250: reshape [-1,2P] -> MoE or affine -> [B,T,V]."""
    B, T, D = enc_out.shape
    x = enc_out.reshape(-1, D)
    if cfg.num_experts > 0:
        y = create_moe(x, p["Variable"], p["Variable_1"], p["Variable_2"], p["Variable_3"],
                       cfg.num_targets, cfg.num_experts, cfg.moe_temp)
    else:
        y = x @ p["Variable"] + p["Variable_1"]
    return y.reshape(B, T, cfg.num_targets)


class _CTC(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, seq_len):
        loss, grad = ctc_loss_grad(logits.detach().numpy(), labels.numpy(), seq_len.numpy())
        ctx.save_for_backward(torch.from_numpy(grad).to(logits.dtype))
        return torch.from_numpy(loss).to(logits.dtype)

    @staticmethod
    def backward(ctx, gout):
        (grad,) = ctx.saved_tensors
        return grad * gout.view(-1, 1, 1), None, None


def ctc_loss_sum(logits, labels, seq_len):
    """graph.py:109-116: per-utt tf.nn.ctc_loss then reduce_sum."""
    return _CTC.apply(logits, labels, seq_len).sum()


def l2_loss(p, weight):
    """graph.py:183-189: sum(v^2)/2 over variables whose NAME lacks 'bias'."""
    tot = 0.0
    for k, v in p.items():
        if "bias" not in k:
            tot = tot + (v * v).sum() / 2
    return tot * weight


def training_loss(p, cfg, nnet_input, seq_len, labels, l2_decay_weight=1e-5):
    """eval_loss (pure CTC sum) and the regularised loss the optimizer sees."""
    enc, _ = blstm_forward(p, cfg, nnet_input, seq_len)
    logits = output_layer(p, cfg, enc)
    ctc = ctc_loss_sum(logits, labels, seq_len)
    return ctc, ctc + l2_loss(p, l2_decay_weight), logits


def clip_by_global_norm(grads: Dict[str, torch.Tensor], clip_norm):
    """tf.clip_by_global_norm: g * clip / max(||g||, clip)."""
    gn = torch.sqrt(sum((g.double() ** 2).sum() for g in grads.values()))
    scale = clip_norm / max(float(gn), clip_norm)
    return {k: g * scale for k, g in grads.items()}, float(gn)


def adam_step(p, grads, state, lr, beta1=0.9, beta2=0.999, eps=1e-8):
    """tf.train.AdamOptimizer (training/adam.py): lr_t = lr*sqrt(1-b2^t)/(1-b1^t);
    var -= lr_t * m / (sqrt(v) + eps)."""
    state["t"] = state.get("t", 0) + 1
    t = state["t"]
    lr_t = lr * math.sqrt(1 - beta2 ** t) / (1 - beta1 ** t)
    for k in p:
        m = state.setdefault("m/" + k, torch.zeros_like(p[k]))
        v = state.setdefault("v/" + k, torch.zeros_like(p[k]))
        m.mul_(beta1).add_(grads[k], alpha=1 - beta1)
        v.mul_(beta2).addcmul_(grads[k], grads[k], value=1 - beta2)
        p[k] = p[k] - lr_t * m / (v.sqrt() + eps)
    return p


def sgd_step(p, grads, state, lr):
    for k in p:
        p[k] = p[k] - lr * grads[k]
    return p


def momentum_step(p, grads, state, lr, momentum=0.9):
    """tf.train.MomentumOptimizer: accum = momentum*accum + g; var -= lr*accum."""
    for k in p:
        a = state.setdefault("a/" + k, torch.zeros_like(p[k]))
        a.mul_(momentum).add_(grads[k])
        p[k] = p[k] - lr * a
    return p


def greedy_decode(logits, seq_len):
    """tf.nn.ctc_greedy_decoder(merge_repeated=True) (graph.py:138-142): argmax, collapse, drop blank."""
    B, T, V = logits.shape
    am = logits.argmax(-1).numpy()
    out = []
    for b in range(B):
        prev, seq = -1, []
        for t in range(int(seq_len[b])):
            c = int(am[b, t])
            if c != prev and c != V - 1:
                seq.append(c)
            prev = c
        out.append(seq)
    return out


def edit_distance(a, b):
    """tf.edit_distance(normalize=False) for one pair (graph.py:143-149)."""
    n, m = len(a), len(b)
    d = list(range(m + 1))
    for i in range(1, n + 1):
        prev, d[0] = d[0], i
        for j in range(1, m + 1):
            cur = d[j]
            d[j] = min(d[j] + 1, d[j - 1] + 1, prev + (a[i - 1] != b[j - 1]))
            prev = cur
    return d[m]



# ------------------------------------------------------------------------------------------------
# create_logits_lstm (nnet/lstm.py:125-368), functional core: uni-directional stack of
# DropoutWrapper([ResidualWrapper](LSTMCell(use_peepholes=True, num_proj))) run by dynamic_rnn(scope="drnn{i}"), default
# forget_bias 1.0, no ResidualWrapper on layer 0 when input_dim != num_projects (lstm.py:236-260), then the affine /
# mixture output layer over [N, num_projects] (lstm.py:318-345).  The builder's feature projection, ornn and
# orthogonality terms call helpers that do not exist in the reference and are not restated.
def lstm_param_order(cfg: OracleConfig) -> List[str]:
    names = []
    for i in range(cfg.num_layers):
        p = "drnn%d/lstm_cell" % i
        names += [p + "/kernel", p + "/bias", p + "/w_f_diag", p + "/w_i_diag", p + "/w_o_diag", p + "/projection/kernel"]
    return names + (["Variable", "Variable_1", "Variable_2", "Variable_3"] if cfg.num_experts > 0 else ["Variable", "Variable_1"])


def init_lstm_params(cfg: OracleConfig, seed=0, dtype=torch.float64, bias_scale=0.0) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    H, P = cfg.num_neurons, cfg.num_projects
    p: Dict[str, torch.Tensor] = {}
    for i in range(cfg.num_layers):
        din = cfg.input_dim if i == 0 else P
        pre = "drnn%d/lstm_cell" % i
        p[pre + "/kernel"] = _glorot((din + P, 4 * H), g, dtype)
        p[pre + "/bias"] = (torch.randn(4 * H, generator=g, dtype=torch.float64) * bias_scale).to(dtype)
        for w in ("w_f_diag", "w_i_diag", "w_o_diag"):
            p[pre + "/" + w] = _glorot((H,), g, dtype)
        p[pre + "/projection/kernel"] = _glorot((H, P), g, dtype)
    std = 1.0 / math.sqrt(P)                                               # lstm.py:331
    if cfg.num_experts > 0:
        K, V = cfg.num_experts, cfg.num_targets
        p["Variable"] = _trunc_normal((P, K), std, g, dtype)
        p["Variable_1"] = (torch.randn(K, generator=g, dtype=torch.float64) * bias_scale).to(dtype)
        p["Variable_2"] = _trunc_normal((P, K * V), std, g, dtype)
        p["Variable_3"] = (torch.randn(K * V, generator=g, dtype=torch.float64) * bias_scale).to(dtype)
    else:
        p["Variable"] = _trunc_normal((P, cfg.num_targets), std, g, dtype)
        p["Variable_1"] = (torch.randn(cfg.num_targets, generator=g, dtype=torch.float64) * bias_scale).to(dtype)
    assert list(p.keys()) == lstm_param_order(cfg)
    return p


def lstm_forward(p, cfg: OracleConfig, nnet_input, seq_len, keep_prob=1.0, masks=None):
    """Returns the last layer's output [B,T,P].  masks: optional {layer: [B,T,P]} 0/1 tensors."""
    x = nnet_input
    for i in range(cfg.num_layers):
        pre = "drnn%d/lstm_cell" % i
        cellp = (p[pre + "/kernel"], p[pre + "/bias"], p[pre + "/w_f_diag"], p[pre + "/w_i_diag"], p[pre + "/w_o_diag"],
                 p[pre + "/projection/kernel"])
        residual = not (i == 0 and cfg.input_dim != cfg.num_projects)      # lstm.py:236
        x, _ = dynamic_rnn(x, seq_len, cellp, 1.0, keep_prob, None if masks is None else masks.get(i), residual=residual)
    return x


def lstm_training_loss(p, cfg, nnet_input, seq_len, labels, l2_decay_weight=1e-5, keep_prob=1.0, masks=None):
    enc = lstm_forward(p, cfg, nnet_input, seq_len, keep_prob, masks)
    logits = output_layer(p, cfg, enc)
    ctc = ctc_loss_sum(logits, labels, seq_len)
    return ctc, ctc + l2_loss(p, l2_decay_weight), logits
