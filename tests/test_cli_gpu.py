"""GPU tests of the callers either side of the path: device splice/subsample vs the host statement of
nnet/tfrecord.py:28-51, and the four command-line drivers (bin/nnet-{init,train,validate,forward}.py of the reference)
end to end on TFRecord files, checking the log lines the shell drivers grep and the Kaldi archive they hand to the decoder."""
import os

import numpy as np
import pytest
import torch

from lstm_ctc_b200 import kaldi_io, tfrecord as tfr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("lc,rc,sub", [(1, 1, 3), (0, 0, 2), (2, 1, 0), (0, 0, 0), (3, 3, 4)])
def test_device_splice_equals_host(cuda_dev, lc, rc, sub):
    rng = np.random.RandomState(lc * 10 + rc + sub)
    B, T, D = 5, 23, 7
    lens = np.array([23, 1, 9, 17, 4], dtype=np.int32)
    x = np.zeros((B, T, D), np.float32)
    for b in range(B):
        x[b, :lens[b]] = rng.randn(lens[b], D)
    d = torch.device("cuda:0")
    out, lo = tfr.splice_subsample_device(torch.tensor(x, device=d), torch.tensor(lens, device=d), lc, rc, sub)
    out, lo = out.cpu().numpy(), lo.cpu().numpy()
    for b in range(B):
        want = tfr.splice_subsample_host(x[b, :lens[b]], lc, rc, sub)
        assert lo[b] == want.shape[0]
        assert np.array_equal(out[b, :lo[b]], want)                  # bit-exact: it is a gather
        assert not out[b, lo[b]:].any()                              # zero padded like padded_batch


def _write_corpus(tmp, n_utts, D, V, seed):
    rng = np.random.RandomState(seed)
    scp = os.path.join(tmp, "feats.scp")
    lens = sorted(rng.randint(30, 61, size=n_utts))
    with open(scp, "w") as fh:
        for i, n in enumerate(lens):
            x = rng.randn(n, D).astype(np.float32)
            y = rng.randint(0, V - 1, size=max(1, n // 12))
            p = os.path.join(tmp, "utt%03d.tfrecords" % i)
            tfr.write_tfrecord(p, x, y)
            fh.write("utt%03d %d %d 1 %s\n" % (i, n, D, p))
    return scp, lens


def test_cli_init_train_validate_forward(cuda_dev, tmp_path, capfd):
    from lstm_ctc_b200 import cli
    tmp = str(tmp_path)
    D, V = 8, 11
    scp, lens = _write_corpus(tmp, 12, D, V, 0)
    cfg = os.path.join(tmp, "nnet.config")
    with open(cfg, "w") as fh:
        fh.write("nnet_type blstm\ninput_dim %d\nleft_context 1\nright_context 1\nsubsample 3\nnum_layers 2\n"
                 "num_neurons 64\nnum_projects 64\nnum_targets %d\nuse_peepholes true\nnum_experts 4\nmoe_temp 10.0\n"
                 "dropout_rate 0.9\n" % (D, V))
    n0, n1 = os.path.join(tmp, "nnet.0"), os.path.join(tmp, "nnet.1")
    cli.nnet_init([scp, cfg, n0, "--objective", "ctc", "--batch-size", "4"])
    err = capfd.readouterr().err
    assert "INFO:tensorflow:cv_loss = " in err
    # the model is a TF checkpoint-V2 bundle under the prefix, as tf.train.Saver leaves it (nnet-init.py:77-79)
    assert os.path.exists(n0 + ".index") and os.path.exists(n0 + ".data-00000-of-00001") and os.path.exists(os.path.join(tmp, "checkpoint"))
    cv0 = float(err.split("cv_loss = ")[1].split()[0])
    for it in range(3):                                               # three "epochs", each a fresh process in the reference
        cli.nnet_train([scp, cfg, n0 if it == 0 else n1, n1, "--objective", "ctc", "--optimizer", "adam", "--learn-rate", "0.004",
                        "--batch-size", "4", "--shuffle", "false", "--device-splice", "true" if it == 1 else "false"])
    err = capfd.readouterr().err
    assert err.count("INFO:tensorflow:tr_loss = ") == 3 and 'saving nnet to "%s"' % n1 in err
    cli.nnet_validate([scp, cfg, n1, "--objective", "ctc", "--batch-size", "4", "--evaluate", "true"])
    err = capfd.readouterr().err
    cv1 = float(err.split("cv_loss = ")[1].split()[0])
    assert "INFO:tensorflow:cv_eval = " in err
    assert np.isfinite(cv0) and np.isfinite(cv1) and cv1 < cv0        # training on the same data lowers its loss
    # inference: log-softmax posteriors as a Kaldi archive; host and device splicing give the same matrices
    arks = []
    for ds in ("false", "true"):
        ark = os.path.join(tmp, "post_%s.ark" % ds)
        cli.nnet_forward([scp, cfg, n1, "ark:" + ark, "--device-splice", ds])
        arks.append(kaldi_io.read_float_matrix_ark(ark))
    capfd.readouterr()
    assert [k for k, _ in arks[0]] == ["utt%03d" % i for i in range(12)]
    for (k, a), (_, b), n in zip(arks[0], arks[1], lens):
        assert a.shape == (n // 3, V)
        assert np.allclose(np.exp(a).sum(1), 1.0, atol=1e-4)
        assert np.array_equal(a, b)               # same inputs bit for bit -> same posteriors bit for bit
    # unsupported objective: fatal log + exit 1 (nnet-train.py:70-77)
    with pytest.raises(SystemExit) as e:
        cli.nnet_validate([scp, cfg, n1, "--batch-size", "4"])
    assert e.value.code == 1 and "unsupported objective: xent" in capfd.readouterr().err


def test_cli_with_nnet_type_lstm(cuda_dev, tmp_path, capfd):
    """The same four CLIs on the uni-directional stack (nnet_type lstm): the bundle holds the uni-directional
    variables under TF's names, training lowers cv_loss, posteriors are normalised, cudnnlstm is refused like the reference's
    unsupported types (nnet-train.py:70-77)."""
    from lstm_ctc_b200 import cli, tf_bundle
    tmp = str(tmp_path)
    D, V = 8, 11
    scp, lens = _write_corpus(tmp, 8, D, V, 1)
    cfg = os.path.join(tmp, "nnet.config")
    with open(cfg, "w") as fh:
        fh.write("nnet_type lstm\ninput_dim %d\nleft_context 1\nright_context 1\nsubsample 3\nnum_layers 3\n"
                 "num_neurons 64\nnum_projects 24\nnum_targets %d\nnum_experts 0\ndropout_rate 0.9\n" % (D, V))   # 3 * 8 = 24 = P: residual on layer 0 too
    n0, n1 = os.path.join(tmp, "nnet.0"), os.path.join(tmp, "nnet.1")
    cli.nnet_init([scp, cfg, n0, "--objective", "ctc", "--batch-size", "4"])
    err = capfd.readouterr().err
    cv0 = float(err.split("cv_loss = ")[1].split()[0])
    names = sorted(tf_bundle.read_bundle(n0))
    assert "drnn0/lstm_cell/kernel" in names and "drnn2/lstm_cell/projection/kernel" in names and not any(k.startswith(("fd", "bd")) for k in names)
    assert tf_bundle.read_bundle(n0)["drnn1/lstm_cell/kernel"].shape == (24 + 24, 4 * 64)
    for it in range(3):
        cli.nnet_train([scp, cfg, n0 if it == 0 else n1, n1, "--objective", "ctc", "--optimizer", "adam", "--learn-rate", "0.004",
                        "--batch-size", "4", "--shuffle", "false"])
    capfd.readouterr()
    cli.nnet_validate([scp, cfg, n1, "--objective", "ctc", "--batch-size", "4"])
    cv1 = float(capfd.readouterr().err.split("cv_loss = ")[1].split()[0])
    assert np.isfinite(cv0) and np.isfinite(cv1) and cv1 < cv0
    ark = os.path.join(tmp, "post.ark")
    cli.nnet_forward([scp, cfg, n1, "ark:" + ark])
    capfd.readouterr()
    for (k, a), n in zip(kaldi_io.read_float_matrix_ark(ark), lens):
        assert a.shape == (n // 3, V) and np.allclose(np.exp(a).sum(1), 1.0, atol=1e-4)
    with open(cfg, "w") as fh:
        fh.write("nnet_type cudnnlstm\ninput_dim %d\nleft_context 1\nright_context 1\nsubsample 3\nnum_layers 1\nnum_neurons 64\nnum_projects 24\nnum_targets %d\ndropout_rate 1.0\n" % (D, V))
    with pytest.raises(SystemExit) as e:
        cli.nnet_validate([scp, cfg, n1, "--objective", "ctc", "--batch-size", "4"])
    assert e.value.code == 1 and "unsupported nnet_type: cudnnlstm" in capfd.readouterr().err


def test_cli_forward_batched_equals_one_by_one(cuda_dev, tmp_path, capfd):
    """nnet-forward --batch-size N (length-bucketed minibatches, forward.py) writes the SAME archive as the reference's
    one-utterance-per-run mode: same keys in scp order, matrices bit for bit; --blank-to-front is the select-feats reorder of
    scripts/decode_ctc_lat.sh:161-163; the graph-API path (create_graph_for_inference + Session.run, what the reference's script
    calls) gives the same posteriors."""
    import lstm_ctc_b200 as nnet
    from lstm_ctc_b200 import cli
    tmp = str(tmp_path)
    D, V = 8, 11
    scp, lens = _write_corpus(tmp, 37, D, V, 3)
    # scp order != length order
    lines = open(scp).read().splitlines()
    rng = np.random.RandomState(0)
    perm = rng.permutation(len(lines))
    with open(scp, "w") as fh:
        fh.write("\n".join(lines[i] for i in perm) + "\n")
    lens = [lens[i] for i in perm]
    cfg = os.path.join(tmp, "nnet.config")
    with open(cfg, "w") as fh:
        fh.write("nnet_type blstm\ninput_dim %d\nleft_context 1\nright_context 1\nsubsample 3\nnum_layers 2\n"
                 "num_neurons 64\nnum_projects 64\nnum_targets %d\nuse_peepholes true\nnum_experts 4\nmoe_temp 10.0\n"
                 "dropout_rate 0.9\n" % (D, V))
    n0 = os.path.join(tmp, "nnet.0")
    cli.nnet_init([scp, cfg, n0, "--objective", "ctc", "--batch-size", "4"])
    prior = os.path.join(tmp, "prior.txt")
    with open(prior, "w") as fh:
        fh.write(" ".join("%d" % c for c in rng.randint(1, 100, size=V)) + "\n")
    arks = {}
    for name, extra in (("b1", []), ("b5", ["--batch-size", "5"]), ("b64", ["--batch-size", "64", "--device-splice", "true"]),
                        ("front", ["--batch-size", "16", "--blank-to-front", "true"])):
        ark = os.path.join(tmp, "post_%s.ark" % name)
        cli.nnet_forward([scp, cfg, n0, "ark,scp:%s,%s.scp" % (ark, ark), "--class-prior", prior] + extra)
        arks[name] = kaldi_io.read_float_matrix_ark(ark)
    capfd.readouterr()
    keys = [os.path.splitext(os.path.basename(l.split()[4]))[0] for l in open(scp)]
    for name in arks:
        assert [k for k, _ in arks[name]] == keys, name                       # scp order
    for (k, a), (_, b), (_, c), (_, f), n in zip(arks["b1"], arks["b5"], arks["b64"], arks["front"], lens):
        assert a.shape == (n // 3, V)
        assert np.array_equal(a, b) and np.array_equal(a, c), k               # independent of the batch an utterance travels in
        assert np.array_equal(f[:, 0], a[:, V - 1]) and np.array_equal(f[:, 1:], a[:, :V - 1])
    # the reference's own call sequence (nnet-forward.py:60-96)
    nc = nnet.parse_config(cfg)
    nc["is_training"] = False
    filename, tfrecord, _ = nnet.dataset_from_tfrecords(tfrecords_scp=scp, left_context=1, right_context=1, subsample=3, shuffle=False)
    init, pipeline = nnet.create_pipeline_sequential(filename=filename, tfrecord=tfrecord)
    graph = nnet.create_graph_for_inference(pipeline=pipeline, nnet_config=nc, smooth_factor=1.0)
    sess = nnet.Session()
    sess.run(init)
    nnet.Saver(nnet.trainable_variables()).restore(sess, n0)
    log_prior = nnet.get_class_prior(prior)
    for k, a in arks["b1"][:5]:
        v = sess.run({"filename": graph["filename"], "nnet_output": graph["nnet_output"]})
        ref = np.log(v["nnet_output"]) - log_prior
        assert os.path.splitext(os.path.basename(v["filename"]))[0] == k
        assert np.abs(ref - a).max() < 1e-5
