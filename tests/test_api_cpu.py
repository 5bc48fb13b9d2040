"""CPU: host logic of the reference-facing API -- config parser, class prior, pipeline batching contract,
epoch-loop bookkeeping / log lines, bucket ordering of the data-parallel reducer (no GPU, no kernels)."""
import io
import os
import sys

import numpy as np
import pytest
import torch

import lstm_ctc_b200 as nnet
from lstm_ctc_b200 import funcs, pipeline


def test_parse_config(tmp_path):
    p = tmp_path / "nnet.config"
    p.write_text("# comment\nnnet_type = blstm\ninput_dim = 120\nmoe_temp = 10.0\nuse_peepholes = true\n"
                 "dropout_rate = 0.9   #keep-prob\nnum_experts = 8\nprior_label_path = /x/y\nuse_bn = False\n\n")
    c = nnet.parse_config(str(p))
    assert c == {"nnet_type": "blstm", "input_dim": 120, "moe_temp": 10.0, "use_peepholes": True, "dropout_rate": 0.9,
                 "num_experts": 8, "prior_label_path": "/x/y", "use_bn": False}
    assert isinstance(c["input_dim"], int) and isinstance(c["moe_temp"], float)
    # reference quirk (config.py:47-50): only tokens that START with '#' are dropped, so a spaced trailing comment
    # becomes the value
    p.write_text("num_layers = 4 # four layers\n")
    assert nnet.parse_config(str(p)) == {"num_layers": "layers"}


def test_class_prior_rotates_blank_to_last(tmp_path):
    p = tmp_path / "label.counts"
    p.write_text("[ 50 30 20 0 ]\n")
    lp = nnet.get_class_prior(str(p))
    assert lp.dtype == np.float32 and lp.shape == (4,)
    assert np.allclose(lp[:2], np.log([0.3, 0.2]), atol=1e-6)
    assert lp[2] == np.float32(-1e10)            # zero count -> floor
    assert np.isclose(lp[3], np.log(0.5), atol=1e-6)   # blank (index 0 in the count file) is last


def test_pipeline_padding_contract():
    utts = [{"nnet_input": np.ones((5, 3), np.float32), "nnet_target": np.array([1, 2], np.int64)},
            {"nnet_input": 2 * np.ones((7, 3), np.float32), "nnet_target": np.array([0], np.int64)},
            {"nnet_input": 3 * np.ones((2, 3), np.float32), "nnet_target": np.array([], np.int64)}]
    init, pipe = nnet.create_pipeline_sequence_batch(utts, input_dim=3, batch_size=2)
    assert set(pipe) == {"nnet_input", "sequence_length", "nnet_target", "target_length"}
    src = pipe["nnet_input"].source
    with pytest.raises(RuntimeError):
        src.next()
    init()
    b = src.next()
    assert b["nnet_input"].shape == (2, 7, 3) and b["nnet_input"].dtype == torch.float32
    assert b["nnet_input"][0, 5:].abs().sum() == 0                       # zero padding (pipeline.py:40)
    assert b["nnet_target"].tolist() == [[1, 2], [0, -1]]                 # -1 padding (pipeline.py:41)
    assert b["sequence_length"].tolist() == [5, 7] and b["sequence_length"].dtype == torch.int32
    assert b["target_length"].tolist() == [2, 1]
    b2 = src.next()                                                        # ragged last batch
    assert b2["nnet_input"].shape == (1, 2, 3) and b2["nnet_target"].tolist() == [[-1]]
    with pytest.raises(nnet.OutOfRangeError):
        src.next()


class _FakeSession:
    def __init__(self, batches):
        self.batches = list(batches)

    def run(self, nodes):
        if not self.batches:
            raise nnet.OutOfRangeError()
        return self.batches.pop(0)


def test_train_loop_running_mean_and_log_lines(capsys):
    graph = {k: k for k in ("size", "train", "summary", "loss", "eval_loss", "sequence_length", "eval")}
    sess = _FakeSession([{"size": 10, "eval_loss": 50.0, "eval": 4.0}, {"size": 30, "eval_loss": 60.0, "eval": 3.0},
                         {"size": 0, "eval_loss": 0.0, "eval": 0.0}])
    assert funcs.train(sess, graph, evaluate=True, report_interval=2) is True
    err = capsys.readouterr().err.splitlines()
    # token-weighted mean of eval_loss/size (funcs.py:48-54): (50 + 60) / 40
    assert "INFO:tensorflow:tr_loss = %f" % (110.0 / 40) in err
    assert any(l.startswith("INFO:tensorflow:step = 2, batch_size = 30, loss = 2.75") and "eval = 0.175" in l for l in err)
    assert "INFO:tensorflow:done" in err


def test_validate_logs_cv_lines_and_nan_exits(capsys):
    graph = {k: k for k in ("size", "loss", "eval_loss", "eval")}
    funcs.validate(_FakeSession([{"size": 4, "eval_loss": 8.0, "eval": 2.0}]), graph, evaluate=True)
    err = capsys.readouterr().err
    assert "INFO:tensorflow:cv_loss = 2.000000" in err and "INFO:tensorflow:cv_eval = 0.500000" in err
    with pytest.raises(SystemExit) as e:
        funcs.validate(_FakeSession([{"size": 4, "eval_loss": float("nan"), "eval": 0.0}]), graph)
    assert e.value.code == 1
    assert "nan loss detected" in capsys.readouterr().err


def test_unsupported_optimizer_and_nnet_type():
    assert nnet.get_optimizer("adagrad", 0.1) is None and nnet.get_optimizer("adam", 0.1)["name"] == "adam"
    assert nnet.get_create_logits("cudnnlstm") is None and nnet.get_create_logits(None) is None
    assert nnet.get_create_logits("blstm") is not None and nnet.get_create_logits("lstm") is not None


def test_param_store_bucket_order_matches_backward():
    """The flat gradient buffer is ordered output layer -> top LSTM layer -> ... -> layer 0, so each
    bucket_ready() slice is contiguous and later ones start where the previous ended."""
    from lstm_ctc_b200.blstm import ParamSpec, ParamStore
    from lstm_ctc_b200.dist import bucket_bounds
    specs = [ParamSpec("out/W", (10, 6), True), ParamSpec("out/b", (10,), True)]
    for i in reversed(range(3)):
        specs += [ParamSpec("L%d/a" % i, (7, 5), True), ParamSpec("L%d/bias" % i, (9,), False)]
    ps = ParamStore(specs, torch.device("cpu"))
    lo0, hi0 = bucket_bounds(ps, ["out/W", "out/b"])
    assert lo0 == 0
    prev = hi0
    for i in reversed(range(3)):
        lo, hi = bucket_bounds(ps, ["L%d/a" % i, "L%d/bias" % i])
        assert lo >= prev - 63 and lo % 64 == 0 and hi > lo
        prev = hi
    assert ps.nodecay_ranges() == [(ps.specs["L%d/bias" % i].offset, ps.specs["L%d/bias" % i].offset + 9) for i in (2, 1, 0)]


def test_flow_control_interlock(monkeypatch):
    """BLSTMEncoder.fwd_flow_control (one recurrence launch that waits in-kernel for GEMMs of another stream) is switched off when
    the process runs under a tool that serialises kernels."""
    from lstm_ctc_b200 import blstm
    monkeypatch.delenv("CUDA_INJECTION64_PATH", raising=False)
    monkeypatch.delenv("CUDA_LAUNCH_BLOCKING", raising=False)
    assert blstm._kernels_serialised() is False
    monkeypatch.setenv("CUDA_LAUNCH_BLOCKING", "1")
    assert blstm._kernels_serialised() is True
    monkeypatch.setenv("CUDA_LAUNCH_BLOCKING", "0")
    assert blstm._kernels_serialised() is False
    monkeypatch.setenv("CUDA_INJECTION64_PATH", "/opt/nvidia/nsight-compute/target/libcuda-injection.so")
    assert blstm._kernels_serialised() is True
