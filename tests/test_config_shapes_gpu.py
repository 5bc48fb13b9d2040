"""GPU parity of the WHOLE hot path at the BASELINE.json model dimensions (C1 / C2 / C3) vs the fp64 oracle.

The reference call chain being matched: /root/reference/nnet/bilstm.py:170-203 (stack), moe.py:43-71 (mixture layer),
graph.py:109-116 (CTC, summed loss), graph.py:190 (gradients).  Same seeded inputs and weights on both sides; with dropout the
kernels' counter-based masks are exported and fed to the oracle (TF's RNG stream cannot be matched).

Compared, per case: logits (live frames), summed CTC loss, every variable's gradient, and d loss / d z_t of EVERY layer (what
the BPTT kernel emits, bf16).  Batch sizes 33 and 64 select both recurrence mappings: 33 -> one 16-utterance group per cluster
with a ragged last group; 64 -> two paired sub-groups per cluster (`lstm_rec_fwd2_kernel<16,2>` / `lstm_rec_bwd3_kernel<16,2>` at
H = 512, the 1-D `lstm_rec_bwd2_kernel` BPTT at H = 320).

Three comparisons per case, all written to gpurun_out/r02_parity_config_shapes.json (committed under profiles/):
  cuda_vs_exact   CUDA path vs the fp64 oracle (oracle/model.py)
  cuda_vs_twin    CUDA path vs the oracle's precision twin (oracle/twin.py: same maths, rounded to 16 bits where the device rounds)
  twin_vs_exact   what the mandated 16-bit operands alone do to the exact model
and two weight regimes:
  default init (forget-gate bias variable 0; with LSTMCell's forget_bias = 5 the cell integrates and the stack's gradient norm
      explodes with T -- rounding ONLY weights and inputs to fp16 moves the oracle's own gradients by 4 % at T = 64 and 45 % at
      T = 128, profiles/r02_oracle_fp16_sensitivity.txt).  Asserted vs the exact oracle at T = 64:
        logits max |err| <= 2.5e-2 * max |logit|, rms <= 5e-3 (observed 1.5e-2 / 2.7e-3 at 5 layers: the 16-bit rounding of each
        layer's output is amplified ~2x per layer above it, profiles/r02_parity_diag_c3.txt), loss 2e-3, per-variable gradients and
        per-layer dz 5e-2 normwise (observed 4.3e-2 / 3.9e-2); and vs the twin: 1e-2 / 2e-3 / 1e-3 / 3e-2 / 3e-2.
      At T = 128 the comparison is reported and bounded RELATIVELY: the CUDA path must sit clearly closer to the twin than
      the twin sits to the exact oracle (gradients: < 0.65 x, observed 0.21 x at C3 and 0.52 x at C1; logits rms < 0.5 x, observed 0.21 x),
      i.e. the deviation is the dynamical amplification of the operand precision, not arithmetic disagreement.
  stable (the same weights with every forget-gate bias variable at -4, i.e. an effective forget bias of 1): no amplification, so
      the exact oracle pins LONG sequences at full config dimensions -- C3 B=64 T=128 keep 0.9, C1 B=33 T=128, C2 B=49 T=192:
        logits max <= 2e-3 (observed 6.5e-4), rms <= 5e-4 (1.2e-4), loss <= 2e-4 (1.8e-5), gradients / dz <= 1.5e-2 (7.3e-3 / 6.3e-3).
Length masking is exact: logits rows of padded frames equal the zero-input output row bit for bit, dz rows of padded frames are 0."""
import json
import os

import pytest
import torch

import oracle
from oracle import twin

pytestmark = pytest.mark.gpu

DIMS = {
    "c1": dict(input_dim=120, num_layers=4, num_neurons=320, num_projects=320, num_targets=72, use_peepholes=True, num_experts=0),
    "c2": dict(input_dim=120, num_layers=4, num_neurons=320, num_projects=320, num_targets=72, use_peepholes=True, num_experts=8),
    "c3": dict(input_dim=120, num_layers=5, num_neurons=512, num_projects=512, num_targets=72, use_peepholes=True, num_experts=8),
}
# (dims, B, T, keep_prob, forget-gate bias variable).  fb = 0: TF's default initialisation (zero biases; with LSTMCell's forget_bias = 5
# the cell is an integrator and the stack's gradients explode with T).  fb = -4: the same weights with the forget-gate block of
# every `bias` variable set to -4, i.e. an effective forget bias of 1 -- a stable model, where the exact oracle pins long sequences.
CASES = [("c3", 33, 64, 1.0, 0.0), ("c2", 64, 64, 0.9, 0.0), ("c1", 64, 64, 0.9, 0.0),
         ("c3", 64, 128, 0.9, -4.0), ("c1", 33, 128, 1.0, -4.0), ("c2", 49, 192, 0.9, -4.0),
         ("c3", 64, 128, 0.9, 0.0), ("c1", 33, 128, 1.0, 0.0)]
TOL_TWIN = {"logits": 1e-2, "logits_rms": 2e-3, "loss": 1e-3, "grad": 3e-2, "dz": 3e-2}      # observed <= 3.6e-3 / 6e-4 / 6e-5 / 1.3e-2 / 1.2e-2
TOL_STABLE = {"logits": 2e-3, "logits_rms": 5e-4, "loss": 2e-4, "grad": 1.5e-2, "dz": 1.5e-2}  # observed <= 6.5e-4 / 1.2e-4 / 1.8e-5 / 7.3e-3 / 6.3e-3
TOL = {"logits": 2.5e-2, "logits_rms": 5e-3, "loss": 2e-3, "grad": 5e-2, "dz": 5e-2}
REPORT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out", "r02_parity_config_shapes.json")


def make_batch(cfg, B, T, seed):
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(B, T, cfg.input_dim, generator=g, dtype=torch.float64)
    lens = torch.randint(int(0.6 * T), T + 1, (B,), generator=g).to(torch.int32)
    lens[0] = T
    lens[B // 2] = max(3, T // 5)                      # one short utterance in the middle of a group
    Lmax = T // 6
    labels = -torch.ones(B, Lmax, dtype=torch.int64)
    for b in range(B):
        x[b, lens[b]:] = 0
        n = max(1, min(Lmax, int(lens[b]) // 8))
        labels[b, :n] = torch.randint(0, cfg.num_targets - 1, (n,), generator=g)
    return x, lens, labels


def nnet_config(cfg, keep):
    return {"nnet_type": "blstm", "input_dim": cfg.input_dim, "left_context": 0, "right_context": 0,
            "num_layers": cfg.num_layers, "num_neurons": cfg.num_neurons, "num_projects": cfg.num_projects,
            "num_targets": cfg.num_targets, "use_peepholes": cfg.use_peepholes, "num_experts": cfg.num_experts,
            "moe_temp": cfg.moe_temp, "dropout_rate": keep}


def _mask(n, keep, seed, dev):
    from lstm_ctc_b200 import _lib
    m = torch.empty(n, dtype=torch.uint8, device=dev)
    _lib.check(_lib.lib().lcb_dropout_mask(_lib.ptr(m), n, keep, seed, _lib.stream_ptr()), "mask")
    return m.cpu().double()


def _unpack_dz(dG, T, B, H, Hp):
    """device dz [T*B, 8Hp] (packed column (unit/8)*32 + gate*8 + unit%8 per direction) -> [2][B, T, 4H] in TF gate-block order"""
    d = dG.float().cpu().double().view(T, B, 2, Hp // 8, 4, 8).permute(2, 1, 0, 4, 3, 5).reshape(2, B, T, 4, Hp)[..., :H]
    return d.reshape(2, B, T, 4 * H)


@pytest.mark.parametrize("dims,B,T,keep,fb", CASES, ids=["%s-B%d-T%d-keep%.1f-fb%g" % c for c in CASES])
def test_whole_path_vs_oracle_at_config_shapes(cuda_dev, dims, B, T, keep, fb):
    from lstm_ctc_b200 import _lib
    from lstm_ctc_b200.model import AcousticModel
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = oracle.OracleConfig(**DIMS[dims])
    params = oracle.init_params(cfg, seed=101, bias_scale=0.1)
    x, lens, labels = make_batch(cfg, B, T, seed=202)
    H, P, K, V, nl = cfg.num_neurons, cfg.num_projects, cfg.num_experts, cfg.num_targets, cfg.num_layers
    if fb:
        for k in params:
            if k.endswith("/bias"):
                params[k][2 * H:3 * H] += fb                          # gate blocks i, j, f, o (rnn_cell_impl.LSTMCell)
    stable = fb < 0

    # ---- CUDA path ----
    m = AcousticModel(nnet_config(cfg, keep), cuda_dev, init=False)
    m.from_tf_dict(params)
    m.enc.debug_dz = {}
    loss_sum, _ = m.loss_and_grad(x.float().to(cuda_dev), lens.to(cuda_dev), labels.to(cuda_dev))
    torch.cuda.synchronize()
    assert _lib.lib().lcb_device_error(1) == 0
    logits = m._out_ws(T, B)["logits"].cpu().double()
    grads = {k: v.cpu().double() for k, v in m.to_tf_dict(grads=True).items()}
    N = T * B

    # ---- oracle, with the kernels' dropout masks ----
    masks, mp, md = None, None, None
    if keep < 1.0:
        masks = {}
        for i in range(nl):
            mk = _mask(N * 2 * P, keep, m.enc.dropout_seed(i), cuda_dev).view(T, B, 2 * P).permute(1, 0, 2)
            masks[(i, "f")] = mk[:, :, :P].contiguous()
            masks[(i, "b")] = oracle.model.reverse_sequence(mk[:, :, P:].contiguous(), lens)
        if K > 0:
            seed = m._out_seed
            mp = _mask(N * K, keep, seed, cuda_dev).view(T, B, K).permute(1, 0, 2).reshape(B * T, K, 1)
            md = _mask(N * K * V, keep, seed ^ 0xD1B54A32D192ED03, cuda_dev).view(T, B, V, K).permute(1, 0, 3, 2).reshape(B * T, K, V)
    live = (torch.arange(T).unsqueeze(0) < lens.unsqueeze(1))
    Hp = m.cfg.Hp
    dz_ours = [_unpack_dz(m.enc.debug_dz[i], T, B, H, Hp) for i in range(nl)]
    for d4 in dz_ours:
        if (~live).any():
            assert d4[:, ~live].abs().max().item() == 0.0                # exact length masking of dz

    def reference(kind):
        """(logits, loss, grads, dz[layer][dir]) of the exact fp64 oracle or of its precision twin"""
        p64 = {k: v.clone().requires_grad_(True) for k, v in params.items()}
        trace = {}
        if kind == "exact":
            enc, _ = oracle.blstm_forward(p64, cfg, x, lens, keep_prob=keep, masks=masks, trace=trace)
            if K > 0:
                y = oracle.create_moe(enc.reshape(-1, 2 * P), p64["Variable"], p64["Variable_1"], p64["Variable_2"], p64["Variable_3"],
                                      V, K, cfg.moe_temp, keep_prob=keep, mask_prior=mp, mask_dec=md).reshape(B, T, V)
            else:
                y = oracle.output_layer(p64, cfg, enc)
        else:
            enc = twin.blstm_forward_twin(p64, cfg, x, lens, keep_prob=keep, masks=masks, trace=trace)
            y = twin.output_layer_twin(p64, cfg, enc, keep_prob=keep, mask_prior=mp, mask_dec=md)
        ctc = oracle.ctc_loss_sum(y, labels, lens)
        ctc.backward()
        dz = []
        for i in range(nl):
            key_f, key_b = ((i, "f"), (i, "b"))
            zf = torch.stack([z.grad if z.grad is not None else torch.zeros_like(z) for z in trace[key_f]], 1)      # [B,T,4H]
            zb = torch.stack([z.grad if z.grad is not None else torch.zeros_like(z) for z in trace[key_b]], 1)
            zb = oracle.model.reverse_sequence(zb, lens)              # the backward cell's own time order -> absolute time
            dz.append([zf * live.unsqueeze(-1), zb * live.unsqueeze(-1)])       # (padded steps carry no gradient)
        return y.detach(), ctc.item(), {k: v.grad for k, v in p64.items()}, dz

    def compare(a, b_):
        """errors of a = (logits, loss, grads, dz) against the reference b_"""
        (la, ca, ga, za), (lb, cb, gb, zb) = a, b_
        scale = lb.abs().max().item()
        e = (la - lb)[live]
        r = {"logits_max_err_over_scale": e.abs().max().item() / scale, "logits_rms_err_over_scale": e.pow(2).mean().sqrt().item() / scale,
             "loss_rel_err": abs(ca - cb) / abs(cb),
             "grad_rel_err": {k: ((ga[k] - gb[k]).norm() / (gb[k].norm() + 1e-30)).item() for k in gb},
             "dz_rel_err": {"L%d/%s" % (i, "fd" if d == 0 else "bd"): ((za[i][d] - zb[i][d]).norm() / (zb[i][d].norm() + 1e-30)).item()
                            for i in range(nl) for d in range(2)}}
        r["grad_rel_err_max"] = max(r["grad_rel_err"].values())
        r["dz_rel_err_max"] = max(r["dz_rel_err"].values())
        return r

    ours = (logits, loss_sum.item(), grads, dz_ours)
    ref_twin, ref_exact = reference("twin"), reference("exact")
    rep = {"case": "%s B=%d T=%d keep=%.1f forget-bias-variable=%g" % (dims, B, T, keep, fb), "tolerance_vs_twin": TOL_TWIN, "tolerance_vs_exact_T<=64": TOL,
           "cuda_vs_twin": compare(ours, ref_twin), "cuda_vs_exact": compare(ours, ref_exact),
           "twin_vs_exact": compare(ref_twin, ref_exact)}
    # padded frames: the output layer applied to a zero encoder row (bilstm.py:237-250 quirk Q5) -- identical rows, bit for bit
    if (~live).any() and keep >= 1.0:
        pad_rows = m._out_ws(T, B)["logits"][(~live).to(cuda_dev)]
        rep["padded_rows_identical"] = bool((pad_rows == pad_rows[0]).all())
        assert rep["padded_rows_identical"]
    try:
        os.makedirs(os.path.dirname(REPORT), exist_ok=True)
        allr = json.load(open(REPORT)) if os.path.exists(REPORT) else {}
        allr[rep["case"]] = rep
        json.dump(allr, open(REPORT, "w"), indent=1)
    except OSError:
        pass
    short = lambda r: {k: (round(v, 6) if isinstance(v, float) else v) for k, v in r.items() if not isinstance(v, dict)}
    print(json.dumps({"case": rep["case"], "cuda_vs_twin": short(rep["cuda_vs_twin"]), "cuda_vs_exact": short(rep["cuda_vs_exact"]),
                      "twin_vs_exact": short(rep["twin_vs_exact"])}))
    # (1) arithmetic agreement, any length: CUDA path vs the precision twin
    ct = rep["cuda_vs_twin"]
    if stable or T <= 64:
        assert ct["logits_max_err_over_scale"] < TOL_TWIN["logits"], ct
        assert ct["logits_rms_err_over_scale"] < TOL_TWIN["logits_rms"], ct
        assert ct["loss_rel_err"] < TOL_TWIN["loss"], ct
        assert ct["grad_rel_err_max"] < TOL_TWIN["grad"], ct["grad_rel_err"]
        assert ct["dz_rel_err_max"] < TOL_TWIN["dz"], ct["dz_rel_err"]
    else:       # default initialisation beyond T = 64: even fp32-level differences are amplified ~1e4-fold; the CUDA path must sit
        #         much closer to the twin than the twin sits to the exact oracle.  Observed ratio of the gradient distances:
        #         0.21 (C3) / 0.52 (C1) with fp16 pre-activations, 0.18 / 0.34 with fp32 ones (profiles/r02_parity_config_shapes*.json):
        #         each rounding stage the device and the twin share is one more place where their last-bit decisions can differ
        #         (fp32 MMA accumulation vs fp64), and this regime amplifies every such difference
        te_ = rep["twin_vs_exact"]
        assert ct["loss_rel_err"] < TOL_TWIN["loss"], ct
        assert ct["grad_rel_err_max"] < 0.65 * te_["grad_rel_err_max"], (ct["grad_rel_err_max"], te_["grad_rel_err_max"])
        assert ct["logits_rms_err_over_scale"] < 0.5 * te_["logits_rms_err_over_scale"], (ct, te_)
    # (2) against the exact fp64 oracle: asserted where the model's own sensitivity to 16-bit operands is still small (T <= 64);
    #     beyond that the CUDA path must stay as close to the exact oracle as the twin does (same dynamical amplification)
    ce, te = rep["cuda_vs_exact"], rep["twin_vs_exact"]
    if stable:
        for key, tk in (("logits_max_err_over_scale", "logits"), ("logits_rms_err_over_scale", "logits_rms"), ("loss_rel_err", "loss"),
                        ("grad_rel_err_max", "grad"), ("dz_rel_err_max", "dz")):
            assert ce[key] < TOL_STABLE[tk], (key, ce[key], ce["grad_rel_err"], ce["dz_rel_err"])
    elif T <= 64:
        assert ce["logits_max_err_over_scale"] < TOL["logits"], ce
        assert ce["logits_rms_err_over_scale"] < TOL["logits_rms"], ce
        assert ce["loss_rel_err"] < TOL["loss"], ce
        assert ce["grad_rel_err_max"] < TOL["grad"], ce["grad_rel_err"]
        assert ce["dz_rel_err_max"] < TOL["dz"], ce["dz_rel_err"]
    else:
        assert ce["loss_rel_err"] < 2 * TOL["loss"], ce
        assert ce["grad_rel_err_max"] < 1.5 * te["grad_rel_err_max"] + TOL_TWIN["grad"], (ce["grad_rel_err_max"], te["grad_rel_err_max"])
        assert ce["logits_rms_err_over_scale"] < 1.5 * te["logits_rms_err_over_scale"] + TOL_TWIN["logits_rms"], (ce, te)


def test_workspace_stays_bounded_over_varying_lengths(cuda_dev):
    """every distinct T used to allocate a full activation set.  50 training steps with random T and B must
    leave the arena at the size of the largest batch (plus its 1/8 growth slack) and torch's allocated memory bounded."""
    from lstm_ctc_b200.model import AcousticModel
    cfg = oracle.OracleConfig(input_dim=40, num_layers=3, num_neurons=128, num_projects=128, num_targets=30, use_peepholes=True,
                              num_experts=4)
    m = AcousticModel(nnet_config(cfg, 0.9), cuda_dev, seed=5)
    g = torch.Generator().manual_seed(9)
    Tmax, Bmax = 160, 24
    seen = []
    for step in range(50):
        T = int(torch.randint(20, Tmax + 1, (1,), generator=g)) if step != 25 else Tmax
        B = int(torch.randint(3, Bmax + 1, (1,), generator=g)) if step != 25 else Bmax
        x = torch.randn(B, T, 40, generator=g)
        lens = torch.randint(T // 2, T + 1, (B,), generator=g).to(torch.int32)
        lens[0] = T
        for b in range(B):
            x[b, lens[b]:] = 0
        labels = torch.randint(0, 29, (B, 4), generator=g)
        m.loss_and_grad(x.to(cuda_dev), lens.to(cuda_dev), labels.to(cuda_dev))
        m.optimizer_step("adam", 1e-4)
        torch.cuda.synchronize()
        seen.append((m.enc.workspace_bytes(), torch.cuda.memory_allocated()))
    # what one (Tmax, Bmax) batch needs, measured on a fresh model
    m2 = AcousticModel(nnet_config(cfg, 0.9), cuda_dev, seed=5)
    x = torch.randn(Bmax, Tmax, 40, generator=g).to(cuda_dev)
    m2.loss_and_grad(x, torch.full((Bmax,), Tmax, dtype=torch.int32, device=cuda_dev), torch.randint(0, 29, (Bmax, 4), generator=g).to(cuda_dev))
    need = m2.enc.workspace_bytes()
    assert seen[-1][0] <= 1.3 * need, (seen[-1][0], need)
    assert max(s[0] for s in seen[26:]) == seen[26][0]              # no growth after the largest batch has been seen
    assert max(s[1] for s in seen[30:]) <= 1.2 * seen[26][1] + (64 << 20)
