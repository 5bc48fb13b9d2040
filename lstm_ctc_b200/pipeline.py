"""Input pipeline contract of the hot path -- mirrors /root/reference/nnet/pipeline.py.

The reference builds `pipeline` dicts of TF tensors fed by tf.data (padded_batch, pipeline.py:35-61);
here a `pipeline` is a dict of handles bound to one batch source that yields, per step,
  nnet_input      [B,Tmax,D] float32, zero padded          (pipeline.py:40)
  nnet_target     [B,Lmax]  int64, -1 padded               (pipeline.py:41)
  sequence_length [B] int32, target_length [B] int32       (pipeline.py:42-43,57-59)
as PINNED host tensors; the Session copies them to the device.  `dataset` is any re-iterable of utterance
dicts {'nnet_input': [T,D] float32, 'nnet_target': [L] int, optional 'filename'}.  TFRecord decoding
(nnet/tfrecord.py) is outside the hot path."""
import queue
import threading

import numpy as np
import torch


class OutOfRangeError(Exception):
    """Stands in for tf.errors.OutOfRangeError: the batch source is exhausted (epoch end)."""


class PipelineTensor:
    def __init__(self, source, key):
        self.source, self.key = source, key

    def __repr__(self):
        return "<pipeline tensor %r>" % self.key


class _BatchSource:
    def __init__(self, dataset, input_dim, batch_size, sequential=False):
        self.dataset, self.input_dim, self.batch_size, self.sequential = dataset, input_dim, batch_size, sequential
        # (left_context, right_context, subsample) still to be applied -- on the device, by the Session -- when the dataset
        # yields raw frames (tfrecord.dataset_from_tfrecords(device_splice=True)); host tensors then have the raw width
        self.device_splice = getattr(dataset, "device_splice", None)
        if self.device_splice is not None and not sequential:
            self.input_dim = getattr(dataset, "raw_input_dim", input_dim)
        self._it = None
        self.pin = torch.cuda.is_available()

    def initialize(self):
        """(Re)start the epoch.  Batches are assembled into pinned memory by a background thread, two ahead
        (the reference's tf.data pipeline prefetches on host threads as well, tfrecord.py:122-123)."""
        self._stop_producer()                                 # a previous epoch's thread may be blocked on its full queue
        self.epoch_id = getattr(self, "epoch_id", 0) + 1      # consumers drop anything they staged from the previous epoch
        self._it = iter(self.dataset)
        self._q = queue.Queue(maxsize=2)
        self._stop = threading.Event()
        self._thread = threading.Thread(target=self._producer, args=(self._it, self._q, self._stop), daemon=True)
        self._thread.start()

    def _stop_producer(self):
        """Ends the producer of the previous epoch: signal it, drain its queue so a blocked put() returns, join it -- so
        re-initialising never leaks a thread holding pinned batches."""
        th = getattr(self, "_thread", None)
        if th is None:
            return
        self._stop.set()
        while th.is_alive():
            try:
                self._q.get_nowait()
            except queue.Empty:
                pass
            th.join(timeout=0.01)
        self._thread = None

    def _producer(self, it, q, stop):
        def put(item):
            while not stop.is_set():
                try:
                    q.put(item, timeout=0.05)
                    return True
                except queue.Full:
                    continue
            return False

        while not stop.is_set():
            try:
                item = self._assemble(it)
            except OutOfRangeError:
                put(None)
                return
            except Exception as e:          # surface data errors on the consumer side
                put(e)
                return
            if not put(item):
                return

    def next(self):
        if self._it is None:
            raise RuntimeError("pipeline not initialised: sess.run(pipeline_initializer) first")
        item = self._q.get()
        if item is None or isinstance(item, Exception):
            self._q.put(item)               # the producer has exited: stay exhausted / keep failing, never block
            if item is None:
                raise OutOfRangeError()
            raise item
        return item

    def _pinned(self, t):
        return t.pin_memory() if self.pin else t

    def _assemble(self, it):
        if self.sequential:
            try:
                utt = next(it)
            except StopIteration:
                raise OutOfRangeError()
            x = torch.as_tensor(np.asarray(utt["nnet_input"], dtype=np.float32))
            return {"filename": utt.get("filename", ""), "nnet_input": self._pinned(x),
                    "sequence_length": int(utt.get("sequence_length", x.shape[0]))}
        utts = []
        for _ in range(self.batch_size):
            try:
                utts.append(next(it))
            except StopIteration:
                break
        if not utts:
            raise OutOfRangeError()
        B = len(utts)
        lens = [int(u.get("sequence_length", np.asarray(u["nnet_input"]).shape[0])) for u in utts]
        tl = [len(u["nnet_target"]) for u in utts]
        T, L = max(lens), max(max(tl), 1)
        x = torch.zeros(B, T, self.input_dim, dtype=torch.float32)
        y = torch.full((B, L), -1, dtype=torch.int64)
        for b, u in enumerate(utts):
            xi = torch.as_tensor(np.asarray(u["nnet_input"], dtype=np.float32))
            x[b, :xi.shape[0]] = xi
            if tl[b]:
                y[b, :tl[b]] = torch.as_tensor(np.asarray(u["nnet_target"], dtype=np.int64))
        return {"nnet_input": self._pinned(x), "nnet_target": self._pinned(y),
                "sequence_length": self._pinned(torch.tensor(lens, dtype=torch.int32)),
                "target_length": self._pinned(torch.tensor(tl, dtype=torch.int32))}


def create_pipeline_sequence_batch(dataset, input_dim, batch_size=64, batch_threads=8, num_epochs=1):
    """-> (initializer, pipeline) for training / validation (pipeline.py:24-63).  `batch_threads` and
    `num_epochs` are accepted and ignored, exactly like the reference (pipeline.py:27-28 never uses them)."""
    src = _BatchSource(dataset, input_dim, batch_size)
    pipeline = {k: PipelineTensor(src, k) for k in ("nnet_input", "sequence_length", "nnet_target", "target_length")}
    return src.initialize, pipeline


def create_pipeline_sequential(filename, tfrecord, num_epochs=1):
    """-> (initializer, pipeline) for inference: one utterance per run (pipeline.py:66-86).
    `filename` and `tfrecord` are parallel re-iterables (the reference zips two tf.data datasets)."""
    class _Zip:
        def __iter__(self_inner):
            for _ in range(num_epochs):
                for f, r in zip(filename, tfrecord):
                    d = dict(r)
                    d["filename"] = f
                    yield d
    zipped = _Zip()
    zipped.device_splice = getattr(tfrecord, "device_splice", None)
    src = _BatchSource(zipped, None, 1, sequential=True)
    pipeline = {k: PipelineTensor(src, k) for k in ("filename", "nnet_input", "sequence_length")}
    return src.initialize, pipeline


class SyntheticDataset:
    """Seeded synthetic utterances of the BASELINE shapes: N(0,1) features (the reference feeds
    CMVN'd fbank), lengths ~ U{ceil(0.8T)..T} sorted ascending, labels ~ U{0..V-2} with L ~ len/8."""

    def __init__(self, num_utts, T, input_dim, num_targets, seed=777, full_length=False, label_ratio=8):
        g = np.random.RandomState(seed)
        lens = np.full(num_utts, T) if full_length else np.sort(g.randint(int(np.ceil(0.8 * T)), T + 1, size=num_utts))
        self.utts = []
        for i in range(num_utts):
            n = int(lens[i])
            L = max(1, n // label_ratio)
            self.utts.append({"nnet_input": g.standard_normal((n, input_dim)).astype(np.float32),
                              "nnet_target": g.randint(0, num_targets - 1, size=L).astype(np.int64),
                              "filename": "utt%05d.tfrecords" % i})

    def __iter__(self):
        return iter(self.utts)

    def __len__(self):
        return len(self.utts)
