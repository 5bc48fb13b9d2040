"""Batched posterior computation behind `bin/nnet-forward.py` (/root/reference/bin/nnet-forward.py:77-96).

The reference runs ONE utterance per `sess.run` (create_pipeline_sequential, pipeline.py:66-86).  The sm_100a recurrence steps
16 utterances per cluster in lockstep, so a single utterance leaves 15/16 of every MMA and all but two clusters idle; this
engine feeds it length-bucketed minibatches instead and keeps the reference's observable behaviour:

  * one output matrix per utterance, [num_frames, num_targets] float32, key = basename of the TFRecord path without extension,
    written in scp order;
  * softmax(smooth_factor * logits) [-> log] [- log class prior], exactly the optional steps of nnet-forward.py:87-91, computed by
    `lcb_posterior` on the device (plus, optionally, the blank -> column 0 reorder that scripts/decode_ctc_lat.sh:161-163 applies
    with `select-feats` before EESEN's latgen-faster);
  * an utterance's posteriors do not depend on the batch it travels in (bit-identical for any --batch-size: the recurrence,
    the GEMMs' K loops and the row-wise output kernels never mix utterances) -- asserted by tests/test_cli_gpu.py.

Bucketing uses the <num-rows> column of the scp (tfrecord.py:64-70): inside a window of `window_batches * batch_size` consecutive
scp lines utterances are sorted by length and cut into minibatches, so padding stays small and memory bounded; results of a window
are written back in scp order.  TFRecord decoding runs on a small thread pool one window ahead of the device."""
import os
import queue
import threading
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import torch

from .decode import softmax_rows
from .tfrecord import read_tfrecord, splice_subsample_device, splice_subsample_host


def read_scp(tfrecords_scp):
    """[(path, num_rows, num_cols)] in scp order: `<utt-id> <num-rows> <num-cols> <has-label> <path>` (tfrecord.py:64-70)."""
    out = []
    for line in open(tfrecords_scp, "r"):
        tok = line.rstrip().split()
        if tok:
            out.append((tok[4], int(tok[1]), int(tok[2])))
    return out


def plan_batches(num_rows, batch_size, window_batches=4):
    """-> list of windows, each a list of minibatches (lists of scp indices, longest utterance first)."""
    n = len(num_rows)
    win = max(1, batch_size * window_batches)
    windows = []
    for w0 in range(0, n, win):
        idx = list(range(w0, min(n, w0 + win)))
        if batch_size > 1:
            idx.sort(key=lambda i: (-num_rows[i], i))
        windows.append([idx[k:k + batch_size] for k in range(0, len(idx), batch_size)])
    return windows


class BatchedForward:
    def __init__(self, model, entries, left_context=0, right_context=0, subsample=0, device_splice=False, batch_size=1,
                 smooth_factor=1.0, apply_softmax=True, apply_log=True, class_prior=None, blank_to_front=False,
                 window_batches=4, io_threads=4):
        self.model, self.entries = model, entries
        self.ctx = (int(left_context or 0), int(right_context or 0), int(subsample or 0))
        self.device_splice = bool(device_splice) and any(self.ctx)
        self.batch_size = max(1, int(batch_size))
        self.smooth, self.apply_softmax, self.apply_log = float(smooth_factor), bool(apply_softmax), bool(apply_log)
        self.prior = None if class_prior is None else torch.as_tensor(np.asarray(class_prior, dtype=np.float32)).to(model.device)
        self.blank_to_front = bool(blank_to_front)
        self.windows = plan_batches([e[1] for e in entries], self.batch_size, window_batches)
        self.io_threads = max(1, int(io_threads))
        self.timing = {"device_s": 0.0, "batches": 0, "frames": 0, "padded_frames": 0}

    # ---- host side: decode one utterance ----
    def _load(self, i):
        path = self.entries[i][0]
        recs = read_tfrecord(path)
        x = np.asarray(recs[0]["nnet_input"], dtype=np.float32)
        if not self.device_splice:
            lc, rc, sub = self.ctx
            x = splice_subsample_host(x, lc, rc, sub)
        return x

    def _producer(self, q, stop):
        pin = torch.cuda.is_available()
        with ThreadPoolExecutor(self.io_threads) as pool:
            for window in self.windows:
                for batch in window:
                    if stop.is_set():
                        return
                    try:
                        xs = list(pool.map(self._load, batch))
                        T = max(x.shape[0] for x in xs)
                        xb = torch.zeros(len(xs), T, xs[0].shape[1], dtype=torch.float32)
                        for b, x in enumerate(xs):
                            xb[b, :x.shape[0]] = torch.from_numpy(x)
                        lens = torch.tensor([x.shape[0] for x in xs], dtype=torch.int32)
                        item = (batch, xb.pin_memory() if pin else xb, lens)
                    except Exception as e:        # surface data errors on the consumer side
                        item = e
                    while not stop.is_set():
                        try:
                            q.put(item, timeout=0.05)
                            break
                        except queue.Full:
                            continue
                    if isinstance(item, Exception):
                        return
                q_item = ("window_end", [i for batch in window for i in batch])
                while not stop.is_set():
                    try:
                        q.put(q_item, timeout=0.05)
                        break
                    except queue.Full:
                        continue
        q.put(None)

    # ---- device side ----
    def _forward(self, xb, lens, host):
        """Enqueues one minibatch: H2D, forward, posterior, D2H into `host` (pinned).  Returns (view of host, frame counts, event that
        completes when the copy has landed, timing events) WITHOUT synchronising: the next minibatch is enqueued while this one
        is written out."""
        m = self.model
        dev = m.device
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        x = xb.to(dev, non_blocking=True)
        ln = lens.to(dev, non_blocking=True)
        lens_out = lens
        if self.device_splice:
            lc, rc, sub = self.ctx
            x, ln = splice_subsample_device(x, ln, lc, rc, sub)
            lens_out = lens // max(sub, 1)
        logits = m.forward_logits(x, ln, training=False, seq_len_host=lens_out)
        if self.apply_softmax:
            out = m.enc._arena.flat("posterior", logits.numel(), torch.float32).view(logits.shape)
            softmax_rows(logits, self.smooth, apply_log=self.apply_log, log_prior=self.prior,
                         blank_to_front=self.blank_to_front, out=out)
        elif self.prior is not None or self.blank_to_front:
            out = logits - self.prior if self.prior is not None else logits.clone()
            if self.blank_to_front:
                out = torch.cat([out[..., -1:], out[..., :-1]], -1)
        else:
            out = logits
        view = host[:out.numel()].view(out.shape)
        view.copy_(out, non_blocking=True)
        e1.record()
        return view, lens_out, e0, e1

    def run(self, write, report=None):
        """write(key, matrix) is called once per utterance, in scp order; report(n) after every utterance.
        Three stages run concurrently: TFRecord decoding + pinned batch assembly (producer thread and its pool), the device
        (this thread only enqueues), and slicing + archive writing (writer thread, fed through two pinned staging buffers)."""
        q = queue.Queue(maxsize=3)
        stop = threading.Event()
        th = threading.Thread(target=self._producer, args=(q, stop), daemon=True)
        th.start()
        wq = queue.Queue(maxsize=4)
        free = queue.Queue()
        err = []
        done = [0]

        def writer():
            pending = {}
            try:
                while True:
                    item = wq.get()
                    if item is None:
                        return
                    if item[0] == "window_end":
                        for i in sorted(item[1]):
                            key, _ = os.path.splitext(os.path.basename(self.entries[i][0]))
                            write(key, pending.pop(i))
                            done[0] += 1
                            if report:
                                report(done[0])
                        continue
                    _, batch, view, lens_out, e0, e1, buf = item
                    e1.synchronize()
                    self.timing["device_s"] += e0.elapsed_time(e1) * 1e-3
                    self.timing["batches"] += 1
                    self.timing["frames"] += int(lens_out.sum())
                    self.timing["padded_frames"] += int(view.shape[0] * view.shape[1])
                    arr = view.numpy()
                    for b, i in enumerate(batch):
                        pending[i] = arr[b, :int(lens_out[b])].copy()      # (the pinned staging buffer is reused)
                    free.put(buf)
            except Exception as e:              # surfaced by the main thread
                err.append(e)
                free.put(None)

        wt = threading.Thread(target=writer, daemon=True)
        wt.start()
        bufs = 0
        try:
            while not err:
                item = q.get()
                if item is None:
                    break
                if isinstance(item, Exception):
                    raise item
                if item[0] == "window_end":
                    wq.put(item)
                    continue
                batch, xb, lens = item
                need = xb.shape[0] * xb.shape[1] * self.model.cfg.V
                buf = None
                if bufs >= 2:
                    buf = free.get()
                    if buf is None:
                        break
                if buf is None or buf.numel() < need:
                    buf = torch.empty(int(need * 1.25) + 1, dtype=torch.float32, pin_memory=True)
                    bufs += 1 if bufs < 2 else 0
                view, lens_out, e0, e1 = self._forward(xb, lens, buf)
                wq.put(("batch", batch, view, lens_out, e0, e1, buf))
        finally:
            stop.set()
            wq.put(None)
            wt.join()
            th.join(timeout=5)
        if err:
            raise err[0]
        return done[0]
