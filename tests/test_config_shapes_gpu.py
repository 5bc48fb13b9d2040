"""GPU parity of the WHOLE hot path at the BASELINE.json model dimensions (C1 / C2 / C3) vs the fp64 oracle.

The reference call chain being matched: /root/reference/nnet/bilstm.py:170-203 (stack), moe.py:43-71 (mixture layer),
graph.py:109-116 (CTC, summed loss), graph.py:190 (gradients).  Same seeded inputs and weights on both sides; with dropout the
kernels' counter-based masks are exported and fed to the oracle (TF's RNG stream cannot be matched).

Compared, per case: logits (live frames), summed CTC loss, every variable's gradient, and d loss / d z_t of EVERY layer (what
the BPTT kernel emits, bf16).  Batch sizes 33 and 64 select both recurrence mappings: 33 -> one 16-utterance group per cluster
with a ragged last group; 64 -> two paired sub-groups per cluster (`lstm_rec_fwd2_kernel<16,2>` / `lstm_rec_bwd3_kernel<16,2>` at
H = 512, the 1-D `lstm_rec_bwd2_kernel` BPTT at H = 320).

Stated tolerances (the observed errors are written to gpurun_out/r02_parity_config_shapes.json and committed under profiles/):
  logits   max |err| over live frames  <= 1e-2 * max |logit|          (fp16 operands, fp32 accumulate, T <= 128 steps)
  loss     |sum - sum_ref|             <= 2e-3 * |sum_ref|
  grads    ||g - g_ref|| / ||g_ref||   <= 5e-2 per variable             (bf16 gradient operands)
  dz       ||dz - dz_ref|| / ||dz_ref|| <= 5e-2 per layer               (bf16 storage)
Length masking is exact: logits rows of padded frames equal the zero-input output row bit for bit, dz rows of padded frames are 0."""
import json
import os

import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

DIMS = {
    "c1": dict(input_dim=120, num_layers=4, num_neurons=320, num_projects=320, num_targets=72, use_peepholes=True, num_experts=0),
    "c2": dict(input_dim=120, num_layers=4, num_neurons=320, num_projects=320, num_targets=72, use_peepholes=True, num_experts=8),
    "c3": dict(input_dim=120, num_layers=5, num_neurons=512, num_projects=512, num_targets=72, use_peepholes=True, num_experts=8),
}
# (dims, B, T, keep_prob)
CASES = [("c3", 33, 64, 1.0), ("c3", 64, 128, 0.9), ("c2", 64, 64, 0.9), ("c1", 33, 128, 1.0), ("c1", 64, 64, 0.9)]
TOL = {"logits": 1e-2, "loss": 2e-3, "grad": 5e-2, "dz": 5e-2}
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


@pytest.mark.parametrize("dims,B,T,keep", CASES, ids=["%s-B%d-T%d-keep%.1f" % c for c in CASES])
def test_whole_path_vs_oracle_at_config_shapes(cuda_dev, dims, B, T, keep):
    from lstm_ctc_b200 import _lib
    from lstm_ctc_b200.model import AcousticModel
    torch.set_num_threads(os.cpu_count() or 1)
    cfg = oracle.OracleConfig(**DIMS[dims])
    params = oracle.init_params(cfg, seed=101, bias_scale=0.1)
    x, lens, labels = make_batch(cfg, B, T, seed=202)
    H, P, K, V, nl = cfg.num_neurons, cfg.num_projects, cfg.num_experts, cfg.num_targets, cfg.num_layers

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
    p64 = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    trace = {}
    enc, _ = oracle.blstm_forward(p64, cfg, x, lens, keep_prob=keep, masks=masks, trace=trace)
    if K > 0:
        y = oracle.create_moe(enc.reshape(-1, 2 * P), p64["Variable"], p64["Variable_1"], p64["Variable_2"], p64["Variable_3"],
                              V, K, cfg.moe_temp, keep_prob=keep, mask_prior=mp, mask_dec=md).reshape(B, T, V)
    else:
        y = oracle.output_layer(p64, cfg, enc)
    ctc = oracle.ctc_loss_sum(y, labels, lens)
    ctc.backward()

    # ---- compare ----
    rep = {"case": "%s B=%d T=%d keep=%.1f" % (dims, B, T, keep), "tolerance": TOL}
    live = (torch.arange(T).unsqueeze(0) < lens.unsqueeze(1))
    scale = y.detach().abs().max().item()
    rep["logits_max_err_over_scale"] = (logits - y.detach())[live].abs().max().item() / scale
    rep["loss_rel_err"] = abs(loss_sum.item() - ctc.item()) / abs(ctc.item())
    rep["grad_rel_err"] = {k: ((v - p64[k].grad).norm() / (p64[k].grad.norm() + 1e-30)).item() for k, v in grads.items()}
    rep["grad_rel_err_max"] = max(rep["grad_rel_err"].values())
    rep["dz_rel_err"] = {}
    Hp = m.cfg.Hp
    for i in range(nl):
        ours = _unpack_dz(m.enc.debug_dz[i], T, B, H, Hp)
        zf = torch.stack([z.grad if z.grad is not None else torch.zeros_like(z) for z in trace[(i, "f")]], 1)      # [B,T,4H]
        zb = torch.stack([z.grad if z.grad is not None else torch.zeros_like(z) for z in trace[(i, "b")]], 1)
        zb = oracle.model.reverse_sequence(zb, lens)                  # the backward cell's own time order -> absolute time
        for d, ref in enumerate((zf, zb)):
            ref = ref * live.unsqueeze(-1)                            # (padded steps carry no gradient)
            rep["dz_rel_err"]["L%d/%s" % (i, "fd" if d == 0 else "bd")] = ((ours[d] - ref).norm() / (ref.norm() + 1e-30)).item()
            assert ours[d][~live].abs().max().item() == 0.0 if (~live).any() else True       # exact masking
    rep["dz_rel_err_max"] = max(rep["dz_rel_err"].values())
    # padded frames: the output layer applied to a zero encoder row (bilstm.py:237-250 quirk Q5) -- identical rows, bit for bit
    if (~live).any():
        pad_rows = m._out_ws(T, B)["logits"][(~live).to(cuda_dev)]
        rep["padded_rows_identical"] = bool((pad_rows == pad_rows[0]).all()) if keep >= 1.0 else None
        if keep >= 1.0:
            assert rep["padded_rows_identical"]
    try:
        os.makedirs(os.path.dirname(REPORT), exist_ok=True)
        allr = json.load(open(REPORT)) if os.path.exists(REPORT) else {}
        allr[rep["case"]] = rep
        json.dump(allr, open(REPORT, "w"), indent=1)
    except OSError:
        pass
    print(json.dumps({k: v for k, v in rep.items() if not isinstance(v, dict) or k == "tolerance"}))
    assert rep["logits_max_err_over_scale"] < TOL["logits"], rep
    assert rep["loss_rel_err"] < TOL["loss"], rep
    assert rep["grad_rel_err_max"] < TOL["grad"], rep["grad_rel_err"]
    assert rep["dz_rel_err_max"] < TOL["dz"], rep["dz_rel_err"]


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
