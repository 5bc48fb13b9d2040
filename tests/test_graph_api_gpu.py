"""GPU: the reference-facing graph API end to end (what bin/nnet-train.py / nnet-validate.py / nnet-forward.py do):
pipeline -> create_graph_for_training_ctc -> Session.run / nnet.train -> Saver -> validation graph (greedy decode +
edit distance) -> inference graph (posterior).  Checks the returned values against the oracle."""
import numpy as np
import pytest
import torch

import oracle

pytestmark = pytest.mark.gpu

CFG = {"nnet_type": "blstm", "input_dim": 20, "left_context": 0, "right_context": 0, "subsample": 0, "num_layers": 2,
       "num_neurons": 64, "num_projects": 32, "num_targets": 10, "use_peepholes": True, "num_experts": 3, "moe_temp": 10.0,
       "dropout_rate": 1.0}


def _data(n=12, seed=5):
    import lstm_ctc_b200 as nnet
    return nnet.SyntheticDataset(n, T=24, input_dim=20, num_targets=10, seed=seed, label_ratio=6)


def test_training_graph_loss_decreases_and_logs(cuda_dev, capsys, tmp_path):
    import lstm_ctc_b200 as nnet
    ds = _data()
    cfg = dict(CFG, is_training=True)
    init, pipe = nnet.create_pipeline_sequence_batch(ds, 20, batch_size=4)
    graph = nnet.create_graph_for_training_ctc(pipe, cfg, learn_rate=5e-3, clip_norm=5.0, optimizer="adam", seed=3)
    assert {"nnet_input", "sequence_length", "logits", "raw_target", "nnet_target", "size", "eval_loss", "loss", "eval",
            "global_step", "summary", "lrate", "train"} <= set(graph)
    sess = nnet.Session()
    losses = []
    for epoch in range(6):
        sess.run(init)
        tot, n = 0.0, 0
        while True:
            try:
                v = sess.run({k: graph[k] for k in ("size", "train", "eval_loss", "loss", "sequence_length")})
            except nnet.OutOfRangeError:
                break
            assert v["size"] == sum(len(u["nnet_target"]) for u in ds.utts[n:n + 4])
            tot += v["eval_loss"]; n += len(v["sequence_length"])
        losses.append(tot)
    assert losses[-1] < 0.8 * losses[0], losses
    # nnet.train: epoch loop + log line the shell driver greps
    sess.run(init)
    assert nnet.train(sess, graph, evaluate=False, report_interval=2) is True
    assert "INFO:tensorflow:tr_loss = " in capsys.readouterr().err
    # checkpoint of trainable variables only, TF names
    path = str(tmp_path / "nnet.1")
    nnet.Saver(nnet.trainable_variables()).save(sess, path)
    from lstm_ctc_b200 import tf_bundle
    import os
    assert os.path.exists(path + ".index") and os.path.exists(path + ".data-00000-of-00001")   # TF checkpoint-V2 bundle under the prefix
    sd = {k: torch.from_numpy(a) for k, a in tf_bundle.read_bundle(path).items()}
    assert "fd0/frnn0/kernel" in sd and sd["fd0/frnn0/kernel"].shape == (20 + 32, 4 * 64) and "Variable_3" in sd
    assert not any(k.startswith(("m/", "v/", "global_step")) for k in sd)

    # validation graph restores the checkpoint; loss / eval match the oracle on the same weights
    cfg_v = dict(CFG, is_training=False)
    init_v, pipe_v = nnet.create_pipeline_sequence_batch(ds, 20, batch_size=6)
    gv = nnet.create_graph_for_validation_ctc(pipe_v, cfg_v)
    nnet.Saver(nnet.trainable_variables()).restore(sess, path)
    sess.run(init_v)
    v = sess.run({k: gv[k] for k in ("size", "loss", "eval_loss", "eval", "logits")})
    ocfg = oracle.OracleConfig.from_nnet_config(CFG)
    p64 = {k: t.double() for k, t in sd.items()}
    B = 6
    T = max(u["nnet_input"].shape[0] for u in ds.utts[:B])
    x = torch.zeros(B, T, 20, dtype=torch.float64)
    lens = torch.zeros(B, dtype=torch.int32)
    L = max(len(u["nnet_target"]) for u in ds.utts[:B])
    y = -torch.ones(B, L, dtype=torch.int64)
    for b, u in enumerate(ds.utts[:B]):
        n = u["nnet_input"].shape[0]
        x[b, :n] = torch.from_numpy(u["nnet_input"]).double(); lens[b] = n
        y[b, :len(u["nnet_target"])] = torch.from_numpy(u["nnet_target"])
    ctc, _, ref_logits = oracle.training_loss(p64, ocfg, x, lens, y, 0.0)
    assert abs(v["eval_loss"] - ctc.item()) < 3e-3 * abs(ctc.item())
    hyp = oracle.greedy_decode(ref_logits.detach(), lens)
    dist = sum(oracle.edit_distance(h, [int(t) for t in y[b] if t >= 0]) for b, h in enumerate(hyp))
    assert abs(v["eval"] - dist) <= 1.0          # argmax ties under fp16 rounding may flip one symbol
    assert nnet.validate(nnet.Session(), gv, evaluate=True) is True   # pipeline already exhausted -> logs only
    err = capsys.readouterr().err
    assert "cv_loss" in err and "cv_eval" in err


def test_inference_graph_posterior(cuda_dev):
    import lstm_ctc_b200 as nnet
    from lstm_ctc_b200.decode import softmax_rows
    ds = _data(3, seed=9)
    cfg = dict(CFG, is_training=False)
    names = [u["filename"] for u in ds.utts]
    init, pipe = nnet.create_pipeline_sequential(names, ds.utts)
    g = nnet.create_graph_for_inference(pipe, cfg, smooth_factor=0.5, seed=11)
    assert set(g) == {"filename", "nnet_input", "sequence_length", "logits", "nnet_output"}
    sess = nnet.Session()
    sess.run(init)
    model = nnet.trainable_variables()
    sd = {k: v.double() for k, v in model.state_dict().items()}
    ocfg = oracle.OracleConfig.from_nnet_config(CFG)
    for i in range(3):
        v = sess.run({"filename": g["filename"], "nnet_output": g["nnet_output"], "logits": g["logits"]})
        assert v["filename"] == names[i]
        x = torch.from_numpy(ds.utts[i]["nnet_input"]).double().unsqueeze(0)
        lens = torch.tensor([x.shape[1]], dtype=torch.int32)
        ref = oracle.output_layer(sd, ocfg, oracle.blstm_forward(sd, ocfg, x, lens)[0])[0]
        post = torch.softmax(0.5 * ref, -1).numpy()
        assert v["nnet_output"].shape == post.shape
        assert np.abs(v["nnet_output"] - post).max() < 2e-2
        assert np.allclose(v["nnet_output"].sum(-1), 1.0, atol=1e-4)
    with pytest.raises(nnet.OutOfRangeError):
        sess.run({"filename": g["filename"]})
    # log-posterior minus prior (nnet-forward.py:87-91)
    lg = torch.randn(5, 10, device=cuda_dev)
    prior = np.log(np.full(10, 0.1, dtype=np.float32))
    out = softmax_rows(lg, 1.0, apply_log=True, log_prior=prior).cpu()
    assert torch.allclose(out, torch.log_softmax(lg.cpu(), -1) - torch.from_numpy(prior), atol=1e-5)


def test_create_moe_signature(cuda_dev):
    """create_moe(lstm_output, output_dim, num_targets, num_experts, moe_temperature, dropout_rate) (moe.py:29-30)."""
    from lstm_ctc_b200.moe import create_moe
    torch.manual_seed(0)
    N, D, V, K = 70, 64, 9, 5
    x = torch.randn(N, D, device=cuda_dev) * 0.5
    Wp, bp = torch.randn(D, K, device=cuda_dev) * 0.2, torch.randn(K, device=cuda_dev) * 0.1
    W, b = torch.randn(D, K * V, device=cuda_dev) * 0.2, torch.randn(K * V, device=cuda_dev) * 0.1
    y = create_moe(x, D, V, K, 10.0, 1.0, W_prior=Wp, b_prior=bp, W=W, b=b)
    ref = oracle.create_moe(x.double().cpu(), Wp.double().cpu(), bp.double().cpu(), W.double().cpu(), b.double().cpu(), V, K, 10.0)
    assert (y.cpu().double() - ref).abs().max().item() < 1e-2 * ref.abs().max().item()
