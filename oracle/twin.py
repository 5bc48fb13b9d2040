"""oracle/twin.py -- TEST INFRASTRUCTURE.  The "precision twin" of the oracle: the SAME mathematics as oracle/model.py
(/root/reference/nnet/bilstm.py:104-273, moe.py:29-72), evaluated in fp64 on the CPU, but with every tensor rounded to 16 bits
at exactly the points where the CUDA path stores or feeds a 16-bit operand:

  forward   features, layer outputs h, the recurrent operand m_t, W_x, W_proj, the output-layer weights, the hoisted pre-activations
            x W_x + b (G_HALF): fp16;
            the folded recurrent weight W' = W_proj * W_h: formed in full precision, then fp16 (the device folds it once per update
            and multiplies m_{t-1} by it, instead of h_{t-1} = m_{t-1} W_proj by W_h);
            accumulation, biases, gate pre-activations, gates, cell state: full precision (fp32 on the device)
  backward  d loss / d z_t (dz) and d loss / d h (the dX of the layer above): bf16, through gradient-rounding hooks

Why it exists.  A randomly initialised peephole BiLSTM with forget bias 5 amplifies perturbations: rounding only the weights and
the input features to fp16 ONCE moves the fp64 oracle's own gradients by 4 % at T = 64, 45 % at T = 128 and 126 % at T = 192
(tests/test_oracle_model_cpu.py::test_sensitivity_to_fp16_rounding, profiles/r02_oracle_fp16_sensitivity.txt).  Against the exact
oracle a 16-bit implementation can therefore only be checked on short sequences; the twin separates ARITHMETIC agreement (CUDA vs
twin: tight at any length, a bug shows up as an O(1) difference) from the DYNAMICAL amplification of the mandated operand
precision (twin vs exact oracle: reported beside it).  Straight-through estimators keep the twin differentiable: its gradient is
the exact gradient of the rounded forward function, with dz / dh additionally rounded to bf16 like the device's."""
import torch

from .model import OracleConfig, _cell_params, reverse_sequence


class _Round(torch.autograd.Function):
    """forward: round to `fwd` dtype (None: identity); backward: round the gradient to `bwd` dtype (None: identity)."""

    @staticmethod
    def forward(ctx, x, fwd, bwd):
        ctx.bwd = bwd
        return x if fwd is None else x.to(fwd).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return (g if ctx.bwd is None else g.to(ctx.bwd).to(g.dtype)), None, None


def q16(x):
    """value stored as fp16 on the device (saturating like the device casts); gradient passes straight through"""
    return _Round.apply(x.clamp(-65504.0, 65504.0), torch.float16, None)


def gq_bf16(x):
    """identity whose incoming gradient is rounded to bf16 (the device stores that gradient as bf16)"""
    return _Round.apply(x, None, torch.bfloat16)


G_HALF = True        # mirror of BLSTMEncoder.g_half (the device default): x W_x + b is rounded to fp16 once before the recurrence adds to it


def _twin_dynamic_rnn(X16, seq_len, cellp, forget_bias, zs=None):
    """One direction of one layer in the device's formulation.  X16 [B,T,Din] already rounded.  Returns m16 [B,T,H] (fp16-rounded
    o*tanh(c), 0 past seq_len) -- the projection h = m W_proj is a bulk product afterwards, as on the device."""
    kernel, bias, w_f, w_i, w_o, proj = cellp
    B, T, din = X16.shape
    H = kernel.shape[1] // 4
    Wx16 = q16(kernel[:din])
    fold16 = q16(proj @ kernel[din:])                        # W' [H, 4H]
    G = X16.reshape(B * T, din) @ Wx16 + bias                # hoisted projection, all frames
    if G_HALF:
        G = q16(G)                                           # the device stores the hoisted pre-activations as fp16 (BLSTMEncoder.g_half)
    G = G.reshape(B, T, 4 * H)
    c = X16.new_zeros(B, H)
    m = X16.new_zeros(B, H)
    outs = []
    for t in range(T):
        z = G[:, t] + m @ fold16
        z = gq_bf16(z)                                       # dz_t leaves the BPTT kernel as bf16 (operand of dX, wgrad and W' dz)
        if zs is not None:
            if z.requires_grad:
                z.retain_grad()
            zs.append(z)
        i, j, f, o = torch.chunk(z, 4, dim=1)
        if w_f is not None:
            c_new = torch.sigmoid(f + forget_bias + w_f * c) * c + torch.sigmoid(i + w_i * c) * torch.tanh(j)
            m_new = torch.sigmoid(o + w_o * c_new) * torch.tanh(c_new)
        else:
            c_new = torch.sigmoid(f + forget_bias) * c + torch.sigmoid(i) * torch.tanh(j)
            m_new = torch.sigmoid(o) * torch.tanh(c_new)
        m_new = q16(m_new)
        live = (t < seq_len).to(X16.dtype).unsqueeze(1)
        outs.append(m_new * live)
        c = live * c_new + (1 - live) * c
        m = live * m_new + (1 - live) * m
    return torch.stack(outs, 1)


def blstm_forward_twin(p, cfg: OracleConfig, nnet_input, seq_len, keep_prob=1.0, masks=None, trace=None):
    """Twin of oracle.blstm_forward.  masks: {(layer, 'f'|'b'): [B,T,P]} in each direction's own time order (as there).
    Returns the encoder output [B,T,2P] (fp16-rounded values)."""
    x16 = q16(nnet_input)
    finput, binput = x16, reverse_sequence(x16, seq_len)
    for i in range(cfg.num_layers):
        outs = []
        for d, inp in (("f", finput), ("b", binput)):
            cellp = _cell_params(p, cfg, i, "fd" if d == "f" else "bd", "frnn" if d == "f" else "brnn")
            m16 = _twin_dynamic_rnn(inp, seq_len, cellp, cfg.forget_bias,
                                    zs=None if trace is None else trace.setdefault((i, d), []))
            h = m16 @ q16(cellp[5])                          # [B,T,P], fp32 accumulate on the device
            if keep_prob < 1.0 and masks is not None:
                h = h * masks[(i, d)] / keep_prob            # DropoutWrapper mask in the GEMM epilogue, before the fp16 store
            outs.append(h)
        cat = torch.cat([outs[0], reverse_sequence(outs[1], seq_len)], 2)
        cat = gq_bf16(q16(cat))                              # layer output stored fp16; its gradient (dX of the layer above) bf16
        if i == 0 and cfg.input_dim == 2 * cfg.num_projects:
            finput = q16(finput + cat)                       # layer-0 residual (bilstm.py:199-200), fp16 add on the device
        else:
            finput = cat
        binput = reverse_sequence(finput, seq_len)
    return finput


def output_layer_twin(p, cfg: OracleConfig, enc16, keep_prob=1.0, mask_prior=None, mask_dec=None):
    """Twin of oracle.output_layer / create_moe: fp16 weights, fp16 encoder rows, full-precision everything else."""
    B, T, D = enc16.shape
    x = enc16.reshape(-1, D)
    if cfg.num_experts > 0:
        K, V = cfg.num_experts, cfg.num_targets
        y_prior = torch.softmax(x @ q16(p["Variable"]) + p["Variable_1"], dim=1).unsqueeze(2)
        if keep_prob < 1.0 and mask_prior is not None:
            y_prior = y_prior * mask_prior / keep_prob
        y_dec = cfg.moe_temp * torch.tanh(x @ q16(p["Variable_2"]) + p["Variable_3"])
        y_dec = y_dec.reshape(-1, K, V)
        if keep_prob < 1.0 and mask_dec is not None:
            y_dec = y_dec * mask_dec / keep_prob
        y = (y_prior * y_dec).sum(1)
    else:
        y = x @ q16(p["Variable"]) + p["Variable_1"]
    return y.reshape(B, T, cfg.num_targets)
