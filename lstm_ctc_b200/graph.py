"""Graph-construction / loss API of the hot path -- host-side mirror of /root/reference/nnet/graph.py.

Same entry points, argument meaning and returned dict keys as the reference:
  get_create_logits, get_optimizer,
  create_graph_for_validation_ctc(pipeline, nnet_config)                     (graph.py:51-162)
  create_graph_for_training_ctc(pipeline, nnet_config, learn_rate, clip_norm, optimizer, l2_decay_weight)  (:165-209)
  create_graph_for_inference(pipeline, nnet_config, smooth_factor)            (:212-241)
The returned values are `Node` handles; `Session.run(nodes)` pulls ONE minibatch from the pipeline and
executes the step on the sm_100a kernels (the equivalent of one sess.run of the reference, funcs.py:40-41).
"""
import os
import sys

import numpy as np
import torch

from . import _lib
from .model import AcousticModel
from .pipeline import OutOfRangeError, PipelineTensor  # noqa: F401
from . import dist as _dist
from . import tf_bundle


class Node:
    def __init__(self, name, graph):
        self.name, self.graph = name, graph

    def __repr__(self):
        return "<node %s>" % self.name


_default_model = [None]


def trainable_variables():
    """Handle on the variables of the most recently built graph (tf.trainable_variables())."""
    return _default_model[0]


class Saver:
    """tf.train.Saver(tf.trainable_variables()): trainable variables ONLY, in the reference's TF names
    (no optimizer slots, no global_step -- nnet-train.py:83-84,95).

    On disk a checkpoint is what TF's Saver writes for the same call: the V2 tensor bundle `<path>.index` +
    `<path>.data-00000-of-00001` (+ the directory's `checkpoint` state file), written and read by `tf_bundle` without
    TensorFlow, so `nnet.$iter` prefixes are interchangeable with the reference's (scripts/train.sh:123-124).  `restore`
    also accepts the `torch.save` dict earlier versions of this package wrote at the bare path."""

    def __init__(self, var_list=None):
        self.model = var_list if var_list is not None else _default_model[0]

    def save(self, sess, path):
        tf_bundle.write_bundle(path, self.model.state_dict())
        return path

    def restore(self, sess, path):
        if tf_bundle.bundle_exists(path):
            sd = {k: torch.from_numpy(v) for k, v in tf_bundle.read_bundle(path).items()}
        elif os.path.isfile(path):
            sd = torch.load(path, map_location="cpu", weights_only=True)
        else:
            raise tf_bundle.BundleError("no checkpoint at %r (neither %s.index nor a state-dict file)" % (path, path))
        self.model.load_state_dict(sd)


def get_create_logits(string):
    """graph.py:24-34.  'blstm' is the builder both recipes use; 'lstm' is the functional core of the reference's stale
    uni-directional builder (lstm.py of this package); 'cudnnlstm' (a cuDNN wrapper returning a bare tensor) is not offered."""
    if not string:
        return None
    if string == "blstm":
        from .bilstm import create_logits_blstm
        return create_logits_blstm
    if string == "lstm":
        from .lstm import create_logits_lstm
        return create_logits_lstm
    return None


def get_optimizer(string, learning_rate, momentum=0.9):
    """graph.py:37-48: 'adam' | 'sgd' | 'momentum' -> (name, hyper-parameters) consumed by lcb_optimizer_step."""
    if string in ("adam", "sgd", "momentum"):
        return {"name": string, "learning_rate": learning_rate, "momentum": momentum}
    return None


def _edit_distance(a, b):
    n, m = len(a), len(b)
    d = list(range(m + 1))
    for i in range(1, n + 1):
        prev, d[0] = d[0], i
        for j in range(1, m + 1):
            cur = d[j]
            d[j] = min(d[j] + 1, d[j - 1] + 1, prev + (a[i - 1] != b[j - 1]))
            prev = cur
    return d[m]


class _Graph:
    def __init__(self, pipeline, nnet_config, mode, seed=None):
        self.mode = mode
        self.source = pipeline["nnet_input"].source
        self.nnet_config = dict(nnet_config)
        nnet_type = nnet_config.get("nnet_type")
        if get_create_logits(nnet_type) is None:
            raise ValueError("unsupported nnet_type: %s" % nnet_type)
        self.model = AcousticModel(self.nnet_config, seed=seed)
        _default_model[0] = self.model
        if seed is not None:                 # --seed (nnet-train.py:141-142, tf.set_random_seed) seeds dropout too; ranks differ
            self.model.enc.set_dropout_seed(seed, _dist.rank())
        self.train_cfg = None
        self.smooth_factor = 1.0
        self.reducer = _dist.GradientAllReducer(self.model.params)
        self.reducer.broadcast_weights()
        self._scal = torch.zeros(4, dtype=torch.float64, device=self.model.device)
        self._staged = None            # (batch, device tensors, copy-done event) of the NEXT step, or the OutOfRangeError to raise
        self._copy_stream = None
        self._staged_epoch = 0
        # early read-back of a training step's scalars: [CTC loss sum, token count, label-smoothing term] leave the device right
        # behind the CTC kernels (own stream, pinned buffer), so Session.run returns the step's loss without draining the backward
        # pass and the update that are still queued -- the next step's launches follow them without a gap
        self._early = None             # (pinned host tensor, copy-done event) of the running step
        self._d2h_stream = None
        self._scal_host = None

    # host -> device staging ------------------------------------------------------
    def _stage(self):
        """Pull the next minibatch from the pipeline and start its host->device copy on the copy stream, so that it overlaps
        the step that is still running (the reference's tf.data pipeline prefetches the same way, tfrecord.py:122-123).
        End of data is remembered and raised by the step that would have consumed the batch."""
        self._staged_epoch = getattr(self.source, "epoch_id", 0)
        try:
            batch = self.source.next()
        except OutOfRangeError as e:
            self._staged = e
            return
        dev = self.model.device
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(self._copy_stream):
            x = batch["nnet_input"].to(dev, non_blocking=True)
            lens = batch["sequence_length"].to(dev, non_blocking=True)
            y = batch["nnet_target"].to(dev, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        self._staged = (batch, x, lens, y, ev)

    # one sess.run ----------------------------------------------------------------
    def step(self, wanted):
        m = self.model
        dev = m.device
        if self.mode == "infer":
            return self._infer(self.source.next(), wanted)           # may raise OutOfRangeError
        if self._staged is None or self._staged_epoch != getattr(self.source, "epoch_id", 0):
            self._stage()                                             # first step, or the pipeline was re-initialised
        staged, self._staged = self._staged, None
        if isinstance(staged, OutOfRangeError):
            self._staged = staged                                     # stay exhausted
            raise staged
        batch, x, lens, y, copied = staged
        torch.cuda.current_stream().wait_event(copied)
        for t in (x, lens, y):
            t.record_stream(torch.cuda.current_stream())
        self._stage()                                                 # next batch's copy runs beside this step
        seq_len_host = batch["sequence_length"]
        if self.source.device_splice is not None:                     # _splice / _subsample of tfrecord.py:28-51, on the device
            from .tfrecord import splice_subsample_device
            lc, rc, sub = self.source.device_splice
            x, lens = splice_subsample_device(x, lens, lc, rc, sub)
            seq_len_host = seq_len_host // max(sub, 1)
            batch = dict(batch, sequence_length=seq_len_host)
        size = int((batch["nnet_target"] != -1).sum())                # graph.py:105-106
        out = {"size": size, "sequence_length": seq_len_host.numpy(), "summary": None,
               "raw_target": batch["nnet_target"].numpy()}
        self._early = None
        if "train" in wanted:
            tc = self.train_cfg
            self.reducer.begin_step()
            early = self._read_back_early if ("eval" not in wanted and dev.type == "cuda") else None
            self._early_size = size
            loss_sum, _ = m.loss_and_grad(x, lens, y, bucket_ready=self.reducer.bucket_ready, seq_len_host=seq_len_host, on_loss=early)
            self.reducer.finish()
            m.optimizer_step(tc["optimizer"], tc["learn_rate"], tc["clip_norm"], tc["l2_decay_weight"])
            out["train"] = None
            logits = m._out_ws(x.shape[1], x.shape[0])["logits"]
        else:
            logits = m.forward_logits(x, lens, training=False, seq_len_host=seq_len_host)
            loss, _ = m.ctc(logits, y, lens)
            loss_sum = loss.sum()
        reg = m.reg_loss if "train" in wanted else m.label_smoothing(logits)
        ev = self._greedy_edit_distance(logits, batch) if "eval" in wanted else 0.0
        if self._early is not None:
            host, copied = self._early
            copied.synchronize()                                      # the read-back of the step's result; backward + update still run
            out["eval_loss"], regv = float(host[0]), float(host[2])
            if self.reducer.world > 1:
                out["size"] = int(round(float(host[1])))
        elif self.reducer.world > 1:
            # every reported scalar is a sum over the GLOBAL batch (graph.py:105-106,116,120-136,150): CTC loss, token count,
            # label-smoothing term and edit distance travel in one all-reduce
            self._scal[0] = loss_sum.double()
            self._scal[1] = float(size)
            self._scal[2] = reg.double().sum() if reg is not None else 0.0
            self._scal[3] = float(ev)
            self.reducer.all_reduce_scalars(self._scal)
            vals = self._scal.tolist()
            out["eval_loss"], out["size"], regv, ev = vals[0], int(round(vals[1])), vals[2], vals[3]
        else:
            out["eval_loss"] = float(loss_sum.item())                 # D2H read of the step's result
            regv = float(reg.item()) if reg is not None else 0.0
        out["loss"] = out["eval_loss"] + regv                         # graph.py:120-136
        out["global_step"] = m.global_step
        if "logits" in wanted:
            out["logits"] = logits.cpu().numpy()
        if "eval" in wanted:
            out["eval"] = ev
        return out

    def _read_back_early(self, loss_sum, reg):
        """on_loss hook of AcousticModel.loss_and_grad: runs between the CTC kernels and the backward pass.  Packs the step's
        scalars, sums them over the ranks (graph.py:105-106,116 are sums over the global batch; every rank issues this
        all-reduce at the same point, before its first gradient bucket) and copies them to pinned host memory on a side stream."""
        main = torch.cuda.current_stream()
        if self._d2h_stream is None:
            self._d2h_stream = torch.cuda.Stream(device=self.model.device)
            self._scal_host = torch.zeros(4, dtype=torch.float64).pin_memory()
            self._scal_ring = [torch.zeros(4, dtype=torch.float64, device=self.model.device) for _ in range(2)]
            self._early_n = 0
        self._early_n += 1
        sc = self._scal_ring[self._early_n & 1]      # two staging buffers in turn: the main stream never waits for the copy stream
        ready = torch.cuda.Event()
        ready.record(main)
        with torch.cuda.stream(self._d2h_stream):                     # (the packing kernels too: nothing is added to the main stream)
            self._d2h_stream.wait_event(ready)
            sc[0] = loss_sum.double()
            sc[1] = float(self._early_size)
            sc[2] = reg.double().sum() if reg is not None else 0.0
            loss_sum.record_stream(self._d2h_stream)
            if reg is not None:
                reg.record_stream(self._d2h_stream)
            self.reducer.all_reduce_scalars(sc)
            self._scal_host.copy_(sc, non_blocking=True)
            copied = torch.cuda.Event()
            copied.record(self._d2h_stream)
        self._early = (self._scal_host, copied)

    def _greedy_edit_distance(self, logits, batch):
        """ctc_greedy_decoder(merge_repeated=True) + edit_distance(normalize=False), summed (graph.py:138-150)."""
        from .decode import greedy_decode
        seqs = greedy_decode(logits, batch["sequence_length"].to(logits.device))
        tot = 0.0
        y = batch["nnet_target"].numpy()
        for b, hyp in enumerate(seqs):
            ref = [int(v) for v in y[b] if v != -1]
            tot += _edit_distance(hyp, ref)
        return tot

    def _infer(self, batch, wanted):
        m = self.model
        x = batch["nnet_input"].to(m.device, non_blocking=True).unsqueeze(0)      # graph.py:227
        lens = torch.tensor([batch["sequence_length"]], dtype=torch.int32, device=m.device)
        if self.source.device_splice is not None:
            from .tfrecord import splice_subsample_device
            lc, rc, sub = self.source.device_splice
            x, lens = splice_subsample_device(x, lens, lc, rc, sub)
            batch = dict(batch, sequence_length=int(batch["sequence_length"]) // max(sub, 1))
        T = x.shape[1]
        logits = m.forward_logits(x, lens, training=False, seq_len_host=[int(batch["sequence_length"])])[0]   # squeeze, graph.py:234
        out = {"filename": batch["filename"], "sequence_length": batch["sequence_length"]}
        if "logits" in wanted:
            out["logits"] = logits.cpu().numpy()
        if "nnet_output" in wanted:
            from .decode import softmax_rows
            out["nnet_output"] = softmax_rows(logits, self.smooth_factor).cpu().numpy()   # graph.py:236
        return out


class Session:
    """Stand-in for tf.Session: run(initializer) starts the pipeline; run(nodes) executes one step."""

    def __init__(self, config=None):
        pass

    def run(self, fetches):
        if callable(fetches):
            fetches()
            return None
        if fetches is None:
            return None
        if isinstance(fetches, dict):
            items = list(fetches.items())
        elif isinstance(fetches, (list, tuple)):
            items = list(enumerate(fetches))
        else:
            items = [(0, fetches)]
        nodes = [n for _, n in items if isinstance(n, Node)]
        if not nodes:
            return None
        g = nodes[0].graph
        values = g.step({n.name for n in nodes})
        res = {k: (values.get(n.name) if isinstance(n, Node) else None) for k, n in items}
        if isinstance(fetches, dict):
            return res
        if isinstance(fetches, (list, tuple)):
            return [res[i] for i in range(len(items))]
        return res[0]


def global_variables_initializer():
    return None          # variables are initialised at graph construction (TF default initialisers)


local_variables_initializer = global_variables_initializer

_VALID_KEYS = ("nnet_input", "sequence_length", "logits", "raw_target", "nnet_target", "size", "eval_loss", "loss",
               "eval", "global_step", "summary")


def create_graph_for_validation_ctc(pipeline, nnet_config, seed=None):
    g = _Graph(pipeline, nnet_config, "valid", seed)
    return {k: Node(k, g) for k in _VALID_KEYS}


def create_graph_for_training_ctc(pipeline, nnet_config, learn_rate, clip_norm=5.0, optimizer="sgd",
                                  l2_decay_weight=1e-5, seed=None):
    if get_optimizer(optimizer, learn_rate) is None:
        raise ValueError("unsupported optimizer: %s" % optimizer)
    g = _Graph(pipeline, nnet_config, "train", seed)
    g.train_cfg = {"learn_rate": float(learn_rate), "clip_norm": float(clip_norm), "optimizer": optimizer,
                   "l2_decay_weight": float(l2_decay_weight)}
    graph = {k: Node(k, g) for k in _VALID_KEYS}
    graph["lrate"] = float(learn_rate)
    graph["train"] = Node("train", g)
    return graph


def create_graph_for_inference(pipeline, nnet_config, smooth_factor=1.0, seed=None):
    cfg = dict(nnet_config)
    g = _Graph(pipeline, cfg, "infer", seed)
    g.smooth_factor = float(smooth_factor)
    return {k: Node(k, g) for k in ("filename", "nnet_input", "sequence_length", "logits", "nnet_output")}
