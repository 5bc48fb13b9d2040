"""TFRecord side of the input pipeline -- TF-free mirror of /root/reference/nnet/tfrecord.py.

  write_tfrecord(filename, nnet_input, nnet_target=None)                 (tfrecord.py:129-156)
  dataset_from_tfrecords(tfrecords_scp, left_context, right_context, subsample, shuffle, seed, num_parallel_calls)
        -> (filename, tfrecord, input_dim)                               (tfrecord.py:54-126)

The reference stores ONE tf.train.SequenceExample per file: feature_lists['nnet_input'] = one FloatList of `num_cols`
floats per frame, feature_lists['nnet_target'] = one single-element Int64List per label (tfrecord.py:136-148), framed as a
TFRecord (u64 length, masked CRC-32C of the length, payload, masked CRC-32C of the payload).  Both are restated here from
the published formats (protobuf wire format of tensorflow/core/example/{example,feature}.proto; record framing of
tensorflow/core/lib/io/record_writer.cc) because TensorFlow is not installable in this image: the byte layout is pinned
by the hand-assembled golden record in tests/golden/ (tests/test_io_cpu.py), not by a TF binary.

Splicing (+-context, edge frames replicated) and subsampling follow `_splice` / `_subsample` (tfrecord.py:28-51):
`splice_subsample_host` is the numpy statement of them; with `device_splice=True` the dataset yields the RAW frames and the
Session applies lcb_splice_subsample on the GPU after the host->device copy (same result, (1+lc+rc)x fewer H2D bytes).
"""
import random
import struct
import sys
import time

import numpy as np

_MASK_DELTA = 0xA282EAD8


def _crc32c(data):
    from . import _lib
    data = bytes(data)
    return int(_lib.lib().lcb_crc32c(data, len(data), 0))


def masked_crc32c(data):
    c = _crc32c(data)
    return (((c >> 15) | (c << 17)) + _MASK_DELTA) & 0xFFFFFFFF


# ---------------------------------------------------------------------------------------------- protobuf wire format
def _varint(n):
    out = bytearray()
    n &= (1 << 64) - 1                      # int64 values are written as 10-byte two's-complement varints when negative
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _read_varint(buf, pos):
    shift, val = 0, 0
    while True:
        b = buf[pos]
        pos += 1
        val |= (b & 0x7F) << shift
        if not (b & 0x80):
            return val, pos
        shift += 7


def _ld(field, payload):                    # length-delimited field
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def _feature_float(row):
    vals = np.asarray(row, dtype="<f4").tobytes()
    return _ld(2, _ld(1, vals))             # Feature.float_list(2) { FloatList.value(1), packed }


def _feature_int64(v):
    return _ld(3, _ld(1, _varint(int(v))))  # Feature.int64_list(3) { Int64List.value(1), packed }


def _feature_list_entry(key, features):
    fl = b"".join(_ld(1, f) for f in features)                 # FeatureList.feature(1), repeated
    return _ld(1, _ld(1, key.encode()) + _ld(2, fl))           # FeatureLists.feature_list(1) map entry {key(1), value(2)}


def serialize_sequence_example(nnet_input, nnet_target=None):
    """SequenceExample(feature_lists=...) exactly as write_tfrecord builds it (tfrecord.py:134-153).  Map entries are
    emitted in key order (nnet_input, nnet_target), as protobuf's deterministic map serialisation does."""
    x = np.asarray(nnet_input, dtype=np.float32)
    entries = _feature_list_entry("nnet_input", [_feature_float(r) for r in x])
    if nnet_target is not None:
        entries += _feature_list_entry("nnet_target", [_feature_int64(v) for v in np.asarray(nnet_target).reshape(-1)])
    return _ld(2, entries)                  # SequenceExample.feature_lists(2)


def write_tfrecord(filename, nnet_input, nnet_target=None):
    payload = serialize_sequence_example(nnet_input, nnet_target)
    head = struct.pack("<Q", len(payload))
    with open(filename, "wb") as f:
        f.write(head)
        f.write(struct.pack("<I", masked_crc32c(head)))
        f.write(payload)
        f.write(struct.pack("<I", masked_crc32c(payload)))


def _fields(buf, lo, hi):
    """Yield (field, wire_type, value_or_(start,end)) of the message in buf[lo:hi]."""
    pos = lo
    while pos < hi:
        tag, pos = _read_varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 2:
            n, pos = _read_varint(buf, pos)
            yield field, wt, (pos, pos + n)
            pos += n
        elif wt == 0:
            v, pos = _read_varint(buf, pos)
            yield field, wt, v
        elif wt == 5:
            yield field, wt, (pos, pos + 4)
            pos += 4
        elif wt == 1:
            yield field, wt, (pos, pos + 8)
            pos += 8
        else:
            raise ValueError("unsupported protobuf wire type %d" % wt)


def _parse_float_feature(buf, lo, hi):
    for f, wt, v in _fields(buf, lo, hi):
        if f == 2 and wt == 2:                                  # float_list
            vals = []
            for f2, wt2, v2 in _fields(buf, v[0], v[1]):
                if f2 == 1 and wt2 == 2:                        # packed
                    vals.append(np.frombuffer(buf, dtype="<f4", count=(v2[1] - v2[0]) // 4, offset=v2[0]))
                elif f2 == 1 and wt2 == 5:                      # unpacked
                    vals.append(np.frombuffer(buf, dtype="<f4", count=1, offset=v2[0]))
            return np.concatenate(vals) if len(vals) != 1 else vals[0]
    return np.zeros(0, dtype=np.float32)


def _packed_float_span(buf, lo, hi):
    """(byte offset, byte count) of the packed float payload if the Feature is exactly one packed FloatList, else None."""
    fs = list(_fields(buf, lo, hi))
    if len(fs) != 1 or fs[0][0] != 2 or fs[0][1] != 2:
        return None
    inner = list(_fields(buf, fs[0][2][0], fs[0][2][1]))
    if len(inner) != 1 or inner[0][0] != 1 or inner[0][1] != 2:
        return None
    return inner[0][2][0], inner[0][2][1] - inner[0][2][0]


def _parse_int64_feature(buf, lo, hi):
    out = []
    for f, wt, v in _fields(buf, lo, hi):
        if f == 3 and wt == 2:                                  # int64_list
            for f2, wt2, v2 in _fields(buf, v[0], v[1]):
                if f2 == 1 and wt2 == 2:
                    p = v2[0]
                    while p < v2[1]:
                        x, p = _read_varint(buf, p)
                        out.append(x - (1 << 64) if x >= (1 << 63) else x)
                elif f2 == 1 and wt2 == 0:
                    out.append(v2 - (1 << 64) if v2 >= (1 << 63) else v2)
    return out


def parse_sequence_example(buf):
    """-> {'nnet_input': [T, D] float32, 'nnet_target': [L] int64} (tf.parse_single_sequence_example with the
    FixedLenSequenceFeature specs of tfrecord.py:96-106), decoded by the library's native parser (`lcb_parse_sequence_example`,
    host code that runs without the GIL on the pipeline's decoding threads).  `parse_sequence_example_py` below is the
    interpreter-level statement of the same decoding that the tests compare it with."""
    import ctypes
    from . import _lib
    L = _lib.lib()
    buf = bytes(buf)
    rows, cols, ny = ctypes.c_longlong(0), ctypes.c_longlong(0), ctypes.c_longlong(0)
    st = L.lcb_parse_sequence_example(buf, len(buf), None, 0, None, 0, ctypes.byref(rows), ctypes.byref(cols), ctypes.byref(ny))
    if st != 0:
        raise ValueError("malformed SequenceExample (status %d)" % st)
    has_y = ny.value >= 0
    x = np.empty((rows.value, cols.value), dtype=np.float32)
    y = np.empty(max(ny.value, 0), dtype=np.int64)
    st = L.lcb_parse_sequence_example(buf, len(buf), x.ctypes.data_as(ctypes.c_void_p), x.size, y.ctypes.data_as(ctypes.c_void_p),
                                      y.size, ctypes.byref(rows), ctypes.byref(cols), ctypes.byref(ny))
    if st != 0:
        raise ValueError("malformed SequenceExample (status %d)" % st)
    return {"nnet_input": x, "nnet_target": y} if has_y else {"nnet_input": x}


def parse_sequence_example_py(buf):
    """The same decoding in Python (test cross-check of the native parser)."""
    buf = bytes(buf)
    out = {}
    for f, wt, v in _fields(buf, 0, len(buf)):
        if f != 2 or wt != 2:
            continue                                            # context features: unused by the reference
        for f1, wt1, v1 in _fields(buf, v[0], v[1]):            # FeatureLists.feature_list entries
            if f1 != 1 or wt1 != 2:
                continue
            key, val = None, None
            for f2, wt2, v2 in _fields(buf, v1[0], v1[1]):
                if f2 == 1:
                    key = buf[v2[0]:v2[1]].decode()
                elif f2 == 2:
                    val = v2
            if key is None or val is None:
                continue
            feats = [v3 for f3, wt3, v3 in _fields(buf, val[0], val[1]) if f3 == 1 and wt3 == 2]
            if key == "nnet_input":
                rows = None
                if feats:
                    # all frames have the same encoded size (fixed num_cols): one strided view instead of T parses
                    span = _packed_float_span(buf, feats[0][0], feats[0][1])
                    stride = feats[1][0] - feats[0][0] if len(feats) > 1 else 0
                    regular = span is not None and len(feats) > 1 and \
                        all(b_ - a_ == feats[0][1] - feats[0][0] for a_, b_ in feats) and \
                        all(feats[i + 1][0] - feats[i][0] == stride for i in range(len(feats) - 1))
                    if regular:
                        raw = np.frombuffer(buf, dtype=np.uint8)
                        rows = np.lib.stride_tricks.as_strided(raw[span[0]:], shape=(len(feats), span[1]),
                                                               strides=(stride, 1)).copy().view("<f4")
                    else:
                        rows = np.stack([_parse_float_feature(buf, a_, b_) for a_, b_ in feats])
                out[key] = np.ascontiguousarray(rows, dtype=np.float32) if rows is not None else np.zeros((0, 0), np.float32)
            elif key == "nnet_target":
                vals = []
                for a, b in feats:
                    vals += _parse_int64_feature(buf, a, b)
                out[key] = np.asarray(vals, dtype=np.int64)
    return out


def read_tfrecord(filename, verify_crc=True):
    """All records of one TFRecord file (the reference writes exactly one per file) as parsed utterance dicts."""
    out = []
    with open(filename, "rb") as f:
        data = f.read()
    pos = 0
    while pos < len(data):
        head = data[pos:pos + 8]
        (n,) = struct.unpack("<Q", head)
        (c_len,) = struct.unpack("<I", data[pos + 8:pos + 12])
        payload = data[pos + 12:pos + 12 + n]
        (c_dat,) = struct.unpack("<I", data[pos + 12 + n:pos + 16 + n])
        if verify_crc and (masked_crc32c(head) != c_len or masked_crc32c(payload) != c_dat):
            raise IOError("corrupted TFRecord (CRC mismatch): %s" % filename)     # tf.errors.DataLossError
        out.append(parse_sequence_example(payload))
        pos += 16 + n
    return out


# ---------------------------------------------------------------------------------------------- splice / subsample
def splice_subsample_host(x, left_context=0, right_context=0, subsample=0):
    """numpy statement of _splice (tfrecord.py:28-40) and _subsample (:43-51)."""
    x = np.asarray(x, dtype=np.float32)
    if left_context or right_context:
        n = x.shape[0]
        padded = np.concatenate([np.repeat(x[:1], left_context, 0), x, np.repeat(x[-1:], right_context, 0)], 0)
        x = np.concatenate([padded[i:i + n] for i in range(left_context + right_context + 1)], 1)
    if subsample:
        x = x[np.arange(x.shape[0] // subsample) * subsample]
    return x


def splice_subsample_device(x, lens, left_context=0, right_context=0, subsample=0):
    """x [B,T,D] f32 cuda (zero padded), lens [B] int32 cuda -> (x' [B,T',D(1+lc+rc)], lens') via lcb_splice_subsample."""
    import torch
    from . import _lib
    B, T, D = x.shape
    f = subsample if subsample and subsample > 1 else 1
    Tout = max(T // f, 1)
    out = torch.empty(B, Tout, D * (1 + left_context + right_context), dtype=torch.float32, device=x.device)
    lens = lens.to(device=x.device, dtype=torch.int32).contiguous()
    lens_out = torch.empty_like(lens)
    _lib.check(_lib.lib().lcb_splice_subsample(_lib.ptr(x.contiguous()), _lib.ptr(lens), _lib.ptr(out), _lib.ptr(lens_out), B, T, D,
                                               int(left_context), int(right_context), int(subsample or 0), Tout, _lib.stream_ptr()),
               "lcb_splice_subsample")
    return out, lens_out


class TFRecordDataset:
    """Re-iterable of utterance dicts, one per scp line, in scp order (or shuffled like tfrecord.py:87-91)."""

    def __init__(self, paths, has_label, left_context, right_context, subsample, device_splice, raw_dim):
        self.paths, self.has_label = paths, has_label
        self.ctx = (int(left_context or 0), int(right_context or 0), int(subsample or 0))
        self.raw_input_dim = raw_dim
        # (lc, rc, subsample) the Session must apply on the device; None: already applied on the host below
        self.device_splice = self.ctx if (device_splice and any(self.ctx)) else None

    def __len__(self):
        return len(self.paths)

    def __iter__(self):
        lc, rc, sub = self.ctx
        for p in self.paths:
            recs = read_tfrecord(p)
            for r in recs:
                x = r["nnet_input"]
                if self.device_splice is None:
                    x = splice_subsample_host(x, lc, rc, sub)
                utt = {"nnet_input": x, "filename": p}
                if self.has_label:
                    utt["nnet_target"] = r.get("nnet_target", np.zeros(0, np.int64))
                yield utt


def dataset_from_tfrecords(tfrecords_scp, left_context=0, right_context=0, subsample=0, shuffle=False, seed=None,
                           num_parallel_calls=32, device_splice=False):
    """scp lines: `<utt-id> <num-rows> <num-cols> <has-label> <path>` (tfrecord.py:64-70).  Returns (filename, tfrecord,
    input_dim) with input_dim already multiplied by the splicing width (tfrecord.py:124), like the reference; the two
    iterables plug into create_pipeline_sequence_batch / create_pipeline_sequential.  `num_parallel_calls` is accepted
    for signature compatibility (decoding runs on the pipeline's prefetch thread)."""
    paths, input_dim, has_label = [], None, None
    for line in open(tfrecords_scp, "r"):
        token = line.rstrip().split()
        if not token:
            continue
        num_cols_, has_label_ = int(token[2]), int(token[3])
        paths.append(token[4])
        if input_dim is None:
            input_dim = num_cols_
        if has_label is None:
            has_label = has_label_
        if input_dim != num_cols_:
            sys.stderr.write("FATAL:tensorflow:inconsistent nnet_input dimension in tfrecords: %d vs. %d\n" % (input_dim, num_cols_))
            sys.exit(1)
        if has_label != has_label_:
            sys.stderr.write("FATAL:tensorflow:inconsistent has_label in tfrecords: %d vs. %d\n" % (has_label, has_label_))
            sys.exit(1)
    if shuffle:
        if seed is None:
            seed = time.time()
        random.seed(seed)
        random.shuffle(paths)
    ds = TFRecordDataset(paths, bool(has_label), left_context, right_context, subsample, device_splice, input_dim)
    width = 1 + int(left_context or 0) + int(right_context or 0)
    return list(paths), ds, (input_dim or 0) * width
