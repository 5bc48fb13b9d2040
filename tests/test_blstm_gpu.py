"""GPU parity: the BiLSTM stack (hoisted tcgen05 projections + cluster-persistent recurrence + BPTT)
vs the fp64 oracle restatement of create_logits_blstm (oracle/model.py).

Stated tolerance (fp16 forward operands, bf16 gradient operands, fp32 accumulate / cell state / master
weights): activations within 1e-2 of max|h| (absolute) -- also after 200 recurrent steps -- and parameter
gradients within 4e-2 normwise relative error per variable.  Length masking is exact:
outputs at t >= seq_len[b] are exactly 0."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu


def make_case(H, P, D, L, B, T, peep, seed=0, full=False):
    cfg = oracle.OracleConfig(input_dim=D, num_layers=L, num_neurons=H, num_projects=P, num_targets=8,
                              use_peepholes=peep, num_experts=0)
    params = oracle.init_params(cfg, seed=seed, bias_scale=0.1)
    g = torch.Generator().manual_seed(seed + 1)
    x = torch.randn(B, T, D, generator=g, dtype=torch.float64)
    lens = torch.full((B,), T, dtype=torch.int32) if full else torch.randint(max(1, T // 2), T + 1, (B,), generator=g).to(torch.int32)
    lens[0] = T
    for b in range(B):
        x[b, lens[b]:] = 0
    return cfg, params, x, lens


def nnet_config(cfg):
    return {"input_dim": cfg.input_dim, "left_context": 0, "right_context": 0, "num_layers": cfg.num_layers,
            "num_neurons": cfg.num_neurons, "num_projects": cfg.num_projects, "num_targets": cfg.num_targets,
            "use_peepholes": cfg.use_peepholes, "num_experts": cfg.num_experts, "dropout_rate": 1.0}


def run_ours(cfg, params, x, lens, R=None):
    from lstm_ctc_b200.blstm import BLSTMEncoder, ModelConfig
    dev = torch.device("cuda:0")
    enc = BLSTMEncoder(ModelConfig(nnet_config(cfg)), dev)
    enc.from_tf_dict(params)
    rt = enc.to_tf_dict()
    for k, v in params.items():
        if k.startswith(("fd", "bd")):
            assert torch.equal(rt[k].cpu().double(), v.float().double()), k      # layout round trip is exact
    B, T, D = x.shape
    out = enc.forward(x.float().to(dev), lens.to(dev), training=True)
    out_bt = out.float().view(T, B, -1).permute(1, 0, 2).contiguous()
    grads = None
    if R is not None:
        enc.params.gflat.zero_()
        dX = R.permute(1, 0, 2).contiguous().view(T * B, -1).to(dev).bfloat16().contiguous()
        enc.last_R = dX
        enc.backward(dX)
        grads = {k: v.cpu().double() for k, v in enc.to_tf_dict(grads=True).items()}
    torch.cuda.synchronize()
    from lstm_ctc_b200 import _lib
    assert _lib.lib().lcb_device_error(1) == 0
    return out_bt.cpu().double(), grads, enc


@pytest.mark.parametrize("H,P,D,L,B,T,peep", [
    (64, 64, 24, 1, 5, 9, False),      # one-CTA cluster, ragged batch < 16
    (128, 64, 40, 2, 16, 12, True),    # 2-CTA cluster, peepholes
    (96, 48, 16, 2, 20, 7, True),      # H padded 96 -> 128, two utterance groups
    (320, 320, 120, 2, 16, 10, True),  # WSJ recipe cell: 5-CTA cluster, 64 units / CTA
    (512, 512, 120, 1, 8, 6, True),    # Libri-shape cell: 16-CTA (non-portable) cluster
    (384, 128, 64, 1, 4, 5, False),    # 12-CTA cluster
    (512, 512, 120, 1, 50, 7, True),   # 32 utterances per cluster (MMA N = 32), ragged second group; 4x4 BPTT decomposition
    (320, 320, 64, 1, 52, 6, True),    # 32 utterances per cluster on a 10-CTA cluster (1-D BPTT kernel)
])
def test_forward_backward_vs_oracle(cuda_dev, H, P, D, L, B, T, peep):
    cfg, params, x, lens = make_case(H, P, D, L, B, T, peep)
    p64 = {k: v.clone().requires_grad_(True) for k, v in params.items()}
    ref, _ = oracle.blstm_forward(p64, cfg, x, lens)
    g = torch.Generator().manual_seed(7)
    R = torch.randn(ref.shape, generator=g, dtype=torch.float64).bfloat16().double()
    (ref * R).sum().backward()
    out, grads, _ = run_ours(cfg, params, x, lens, R)
    scale = ref.abs().max().item()
    err = (out - ref.detach()).abs().max().item()
    assert err < 1e-2 * scale, ("activations", err, scale)
    mask = torch.arange(T).unsqueeze(0) >= lens.unsqueeze(1)
    assert out[mask].abs().max().item() == 0.0 if mask.any() else True      # exact length masking
    worst = {}
    for k, v in grads.items():
        rg = p64[k].grad
        rel = ((v - rg).norm() / (rg.norm() + 1e-12)).item()
        worst[k] = rel
    bad = {k: r for k, r in worst.items() if r > 4e-2}
    assert not bad, bad


def test_long_sequence_state_precision(cuda_dev):
    """T=200: the fp32 cell state kept in registers must not drift (bf16 only touches m and the weights)."""
    cfg, params, x, lens = make_case(128, 128, 40, 1, 16, 200, True, seed=3)
    ref, _ = oracle.blstm_forward(params, cfg, x, lens)
    out, _, _ = run_ours(cfg, params, x, lens)
    assert (out - ref).abs().max().item() < 1e-2 * ref.abs().max().item()


def test_final_state_encoder(cuda_dev):
    cfg, params, x, lens = make_case(64, 32, 16, 2, 6, 11, True, seed=5)
    _, enc_ref = oracle.blstm_forward(params, cfg, x, lens)
    _, _, enc = run_ours(cfg, params, x, lens)
    e = enc.encoder_state().cpu().double()
    assert e.shape == enc_ref.shape
    assert (e - enc_ref).abs().max().item() < 3e-2 * enc_ref.abs().max().item()


@pytest.mark.parametrize("H,B,T", [(512, 40, 64), (128, 16, 70), (320, 52, 60)])
def test_split_recurrence_equals_whole(cuda_dev, H, B, T):
    """lcb_lstm_rec_fwd_range: projecting the first scan steps, starting the recurrence on them and resuming it after the
    rest of the projection (BLSTMEncoder.head_frac) gives the same activations, saved states and final states as one launch,
    with ragged lengths (utterances that end / start inside either part) -- and still matches the oracle."""
    from lstm_ctc_b200.blstm import BLSTMEncoder, ModelConfig
    cfg, params, x, lens = make_case(H, H, 24, 2, B, T, True, seed=11)
    lens[1] = 3                                   # ends inside the head part of the forward scan, starts in the tail of the backward one
    x[1, 3:] = 0
    dev = torch.device("cuda:0")
    outs = []
    for fracs in ([], [0.3], [0.25, 0.5, 0.75]):
        enc = BLSTMEncoder(ModelConfig(nnet_config(cfg)), dev)
        enc.from_tf_dict(params)
        enc.fwd_flow_control = False               # consecutive range launches (the flow-controlled single launch: next test)
        enc.head_fracs = fracs
        enc.head_fracs_tight = fracs
        out = enc.forward(x.float().to(dev), lens.to(dev), training=True).float().clone()
        ws = enc._workspace(T, B, True)
        outs.append((out, [m.float().clone() for m in ws["M"]], [c.clone() for c in ws["cst"]], enc.encoder_state().clone()))
    torch.cuda.synchronize()
    from lstm_ctc_b200 import _lib
    assert _lib.lib().lcb_device_error(1) == 0
    (o0, m0, c0, e0), (o1, m1, c1, e1), (o2, m2, c2, e2) = outs
    assert torch.equal(o0, o2) and torch.equal(e0, e2) and all(torch.equal(a, b) for a, b in zip(m0, m2))     # four launches
    valid = (torch.arange(T).unsqueeze(1) < lens.unsqueeze(0)).reshape(T * B).to(dev)      # rows n = t*B + b of live frames
    # bit-identical: the split changes launches, not arithmetic (each MMA issuer thread owns its accumulator, so the sum
    # order inside a time step is fixed)
    assert torch.equal(o0, o1)
    for a, b in zip(m0, m1):
        assert torch.equal(a, b)
    for a, b in zip(c0, c1):
        assert torch.equal(a[valid], b[valid])
    assert torch.equal(e0, e1)
    ref, _ = oracle.blstm_forward(params, cfg, x, lens)
    out_bt = o1.view(T, B, -1).permute(1, 0, 2).cpu().double()
    assert (out_bt - ref).abs().max().item() < 1e-2 * ref.abs().max().item()


@pytest.mark.parametrize("H,B,T,order", [(512, 64, 96, "sorted"), (512, 64, 70, "shuffled"), (512, 80, 64, "sorted"), (512, 33, 50, "shuffled"),
                                         (320, 52, 60, "sorted"), (512, 128, 40, "sorted")])
def test_host_lengths_skip_dead_steps(cuda_dev, H, B, T, order):
    """lcb_lstm_rec_fwd_range_hl: with the lengths given on the host, a 16-utterance group skips the scan steps in which none of
    its utterances is live (and B = 64 / 80 run on 2 x 3 clusters, the surplus groups paired into the first ones).  Activations,
    zero rows past sequence_length, saved states of live frames, final states and the gradients BPTT forms from them are
    bit-identical to the path without host lengths -- in one launch and in several, with groups that retire early, start
    late, or are not live at all in a launch's range."""
    from lstm_ctc_b200.blstm import BLSTMEncoder, ModelConfig
    cfg, params, x, lens = make_case(H, H, 24, 2, B, T, True, seed=23)
    g = torch.Generator().manual_seed(5)
    lens = torch.randint(max(1, T // 4), T + 1, (B,), generator=g, dtype=torch.int32)
    lens[0] = 2                                    # a whole group of very short utterances: not live in most launch ranges
    lens[1:16] = torch.randint(1, 6, (15,), generator=g, dtype=torch.int32)
    if order == "sorted":
        lens = torch.sort(lens).values
    lens[-1] = T
    for b in range(B):
        x[b, int(lens[b]):] = 0
    dev = torch.device("cuda:0")
    dX = torch.randn(T * B, 2 * H, generator=g).to(dev).bfloat16()
    outs = []
    # (launch structure, host lengths, flow control): range launches vs ONE launch whose prefetch warps wait for the projection
    # chunks still running beside it (ready_steps)
    # The flow-controlled launch also publishes its progress (lcb_lstm_rec_fwd_range_pg): the output projection of the first
    # fwd_hproj_fracs of the scan runs beside the rest of it (same GEMM per row: bit-identical).
    cases = (([], None, False, [0.7]), ([], lens, False, [0.7]), ([0.3, 0.6], None, False, [0.7]), ([0.3, 0.6], lens.numpy(), False, [0.7]),
             ([0.5], lens.tolist(), False, [0.7]), ([0.25, 0.5, 0.8], lens, True, [0.7]), ([0.2], None, True, [0.7]),
             ([0.25, 0.5, 0.8], lens, True, [0.3, 0.55, 0.8]), ([0.3], None, True, [0.5]), ([0.25, 0.5], lens, True, []))
    for fracs, host, flow, hfr in cases:
        enc = BLSTMEncoder(ModelConfig(nnet_config(cfg)), dev)
        enc.from_tf_dict(params)
        enc.fwd_flow_control = flow
        enc.fwd_hproj_fracs = hfr
        enc.flow_fracs = fracs
        enc.head_fracs = fracs
        enc.head_fracs_tight = fracs
        out = enc.forward(x.float().to(dev), lens.to(dev), training=True, seq_len_host=host).float().clone()
        ws = enc._workspace(T, B, True)
        saved = ([m.clone() for m in ws["M"]], [c.clone() for c in ws["cst"]], [q.clone() for q in ws["gates"]])
        enc.params.gflat.zero_()
        enc.backward(dX.clone())
        outs.append((out, saved, enc.encoder_state().clone(), enc.params.gflat.clone()))
    torch.cuda.synchronize()
    from lstm_ctc_b200 import _lib
    assert _lib.lib().lcb_device_error(1) == 0
    valid = (torch.arange(T).unsqueeze(1) < lens.unsqueeze(0)).reshape(T * B).to(dev)      # rows n = t*B + b of live frames
    o0, (m0, c0, q0), e0, g0 = outs[0]
    assert o0[~valid].abs().max().item() == 0.0
    for o, (m, c, q), e, gr in outs[1:]:
        assert torch.equal(o, o0) and torch.equal(e, e0)
        for a, b in zip(m, m0):
            assert torch.equal(a, b)                                   # m: every row (zeros past the end)
        for a, b in zip(c, c0):
            assert torch.equal(a[valid], b[valid])
        for a, b in zip(q, q0):
            assert torch.equal(a[valid], b[valid])
        # BPTT reads saved activations of live frames only; its sums run in one fixed order per launch structure
        assert (gr - g0).abs().max().item() <= 1e-5 * g0.abs().max().item()
    ref, _ = oracle.blstm_forward(params, cfg, x, lens)
    out_bt = o0.view(T, B, -1).permute(1, 0, 2).cpu().double()
    assert (out_bt - ref).abs().max().item() < 1e-2 * ref.abs().max().item()


@pytest.mark.parametrize("B,T,frac", [(40, 64, 0.3), (64, 50, 0.5), (16, 33, 0.1)])
def test_split_bptt_equals_whole(cuda_dev, B, T, frac):
    """lcb_lstm_rec_bwd_range: BPTT over two consecutive scan ranges joined by the carry buffer (recurrent dm of the next step,
    carried dc) gives the same dz (bit for bit), input gradient and parameter gradients as one launch; ragged lengths, one
    utterance ending inside the first range of the backward direction's scan."""
    from lstm_ctc_b200 import _lib
    from lstm_ctc_b200.blstm import BLSTMEncoder, ModelConfig
    H = 512
    assert _lib.lib().lcb_lstm_rec_bwd_can_split(H) == 1
    cfg, params, x, lens = make_case(H, H, 24, 2, B, T, True, seed=13)
    lens[1] = 3
    x[1, 3:] = 0
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(4)
    dtop = (torch.randn(T * B, 2 * H, generator=g) * 0.1).to(dev).bfloat16()
    res = []
    for f in (0.0, frac):
        enc = BLSTMEncoder(ModelConfig(nnet_config(cfg)), dev)
        enc.from_tf_dict(params)
        enc.bwd_split_frac = f
        enc.bwd_early_fracs = []
        enc.forward(x.float().to(dev), lens.to(dev), training=True)
        enc.params.gflat.zero_()
        enc.backward(dtop.clone())
        ws = enc._workspace(T, B, True)
        torch.cuda.synchronize()
        res.append(([d.clone() for d in ws["dG"]], enc.params.gflat.clone()))
    assert _lib.lib().lcb_device_error(1) == 0
    (dg0, g0), (dg1, g1) = res
    for a, b in zip(dg0, dg1):
        assert torch.equal(a, b)
    # parameter gradients: same dz, but bias / peephole sums are atomically accumulated per launch and the weight-gradient
    # GEMMs use split-K reduce-adds -> equal up to fp32 summation order
    assert (g0 - g1).abs().max().item() <= 1e-5 * g0.abs().max().item()
    assert g0.abs().max().item() > 0


@pytest.mark.parametrize("B,T,fracs,keep,layers", [(40, 96, [0.75], 0.8, 3), (64, 80, [0.6], 1.0, 3), (24, 130, [0.65, 0.85], 0.9, 3),
                                                   (16, 96, [0.67, 0.85], 0.9, 1), (33, 70, [0.67, 0.85], 1.0, 2)])
def test_early_rows_schedule_equals_serial(cuda_dev, B, T, fracs, keep, layers):
    """backward() with bwd_early_fracs: BPTT of layers 1.. in two or three launches, the rows of dX (fused dropout mask addressed by
    absolute element index) and of the next layer's dM whose dG is final after the first launch computed on a side stream
    beside the second -- three layers so that the three dX buffers rotate.  Same dz bit for bit in every layer that is still
    held (layer 0's depends on every dX above it), parameter gradients equal up to fp32 summation order."""
    from lstm_ctc_b200 import _lib
    from lstm_ctc_b200.blstm import BLSTMEncoder, ModelConfig
    H = 512
    assert _lib.lib().lcb_lstm_rec_bwd_can_split(H) == 1
    cfg, params, x, lens = make_case(H, H, 24, layers, B, T, True, seed=21)      # one layer: only layer 0's released frames
    lens[1] = 5
    x[1, 5:] = 0
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(5)
    dtop = (torch.randn(T * B, 2 * H, generator=g) * 0.1).to(dev).bfloat16()
    nc = nnet_config(cfg)
    nc["dropout_rate"] = keep
    res = []
    # serial | range launches at the release points (round 1) | ONE launch publishing its progress, consumers wait on their stream
    for f, pg in (([], False), (fracs, False), (fracs, True)):
        enc = BLSTMEncoder(ModelConfig(nc), dev)
        enc.from_tf_dict(params)
        enc.bwd_early_fracs = f
        enc.bwd_progress = pg
        enc.forward(x.float().to(dev), lens.to(dev), training=True)
        enc.params.gflat.zero_()
        enc.backward(dtop.clone())
        ws = enc._workspace(T, B, True)
        torch.cuda.synchronize()
        res.append(([d.clone() for d in ws["dG"]], enc.params.gflat.clone(), ws["dX"][1].clone()))
    assert _lib.lib().lcb_device_error(1) == 0
    dg0, g0, dx0 = res[0]
    for dg1, g1, dx1 in res[1:]:
        if layers > 1:
            assert torch.equal(dx0, dx1)               # layer 1's dX (= layer 0's dH), all rows
        for a, b in list(zip(dg0, dg1))[:min(layers, 2)]:      # (a one-layer stack never writes the second dz buffer)
            assert torch.equal(a, b)
        assert (g0 - g1).abs().max().item() <= 1e-5 * g0.abs().max().item()
    assert g0.abs().max().item() > 0


def test_bf16_twins_of_the_saved_activations(cuda_dev):
    """The wgrad operands: the recurrence kernel writes m_t as fp16 AND bf16 (Mout_bf16 of lcb_lstm_rec_fwd_range_pg), the output
    projection writes h as fp16 AND bf16 (lcb_gemm16_twin).  Each twin is the bf16 rounding of the same fp32 value the fp16 row was
    rounded from: they agree to the coarser format's half ulp, zero rows (frames past sequence_length) are zero in both, and the
    gradients backward() forms from them match the ones from converted copies of the fp16 rows."""
    from lstm_ctc_b200 import blstm
    from lstm_ctc_b200.blstm import BLSTMEncoder, ModelConfig
    cfg, params, x, lens = make_case(512, 512, 24, 2, 40, 48, True, seed=31)
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(9)
    B, T = 40, 48
    dX = torch.randn(T * B, 2 * 512, generator=g).to(dev).bfloat16()
    enc = BLSTMEncoder(ModelConfig(nnet_config(cfg)), dev)
    enc.from_tf_dict(params)
    enc.forward(x.float().to(dev), lens.to(dev), training=True)
    ws = enc._workspace(T, B, True)
    valid = (torch.arange(T).unsqueeze(1) < lens.unsqueeze(0)).reshape(T * B).to(dev)
    for i in range(2):
        for f16, bf in ((ws["M"][i], ws["Mbf"][i]), (ws["Hout"][i], ws["Hbf"][i])):
            a, b = f16.float(), bf.float()
            assert (b[~valid] == 0).all() and (a[~valid] == 0).all()
            assert ((a - b).abs() <= 2.0 ** -8 * a.abs() + 1e-7).all()
    enc.params.gflat.zero_()
    enc.backward(dX.clone())
    g_twin = enc.params.gflat.clone()
    # the same backward from converted copies of the fp16 rows (what backward() did before the twins existed)
    for i in range(2):
        ws["Mbf"][i].copy_(ws["M"][i])
        ws["Hbf"][i].copy_(ws["Hout"][i])
    enc.params.gflat.zero_()
    enc.backward(dX.clone())
    g_conv = enc.params.gflat.clone()
    torch.cuda.synchronize()
    assert (g_twin - g_conv).norm().item() <= 2e-3 * g_conv.norm().item()
