"""CPU tests of the data formats either side of the path: TFRecord SequenceExample writer/reader
(nnet/tfrecord.py), splice/subsample semantics, Kaldi float-matrix archives (pyKaldiIO), CRC-32C.  TensorFlow cannot be
installed here, so the byte layout is pinned by vectors assembled by hand from the published formats (protobuf wire
format of example.proto / feature.proto; TFRecord framing; RFC 3720 CRC-32C test vectors; Kaldi binary matrix header)."""
import os
import struct

import numpy as np
import pytest

from lstm_ctc_b200 import _lib, kaldi_io, tfrecord as tfr


def crc(b):
    return int(_lib.lib().lcb_crc32c(bytes(b), len(b), 0))


def test_crc32c_rfc3720_vectors():
    assert crc(b"123456789") == 0xE3069283
    assert crc(b"\x00" * 32) == 0x8A9136AA
    assert crc(b"\xff" * 32) == 0x62A8AB43
    assert crc(bytes(range(32))) == 0x46DD794E
    assert crc(bytes(range(31, -1, -1))) == 0x113FDB5C
    # continuation and unaligned starts agree with the one-shot value
    data = bytes((i * 7 + 3) & 0xFF for i in range(1000))
    for cut in (0, 1, 7, 8, 9, 500, 999, 1000):
        part = int(_lib.lib().lcb_crc32c(data[:cut], cut, 0))
        assert int(_lib.lib().lcb_crc32c(data[cut:], len(data) - cut, part)) == crc(data)


def test_sequence_example_bytes_hand_assembled():
    """x = [[1.0, 2.0]], target = [3]: every byte written out from the .proto field numbers."""
    floats = struct.pack("<2f", 1.0, 2.0)
    feature_f = b"\x12\x0a" + b"\x0a\x08" + floats                      # Feature{float_list(2){value(1) packed}}
    flist_f = b"\x0a" + bytes([len(feature_f)]) + feature_f             # FeatureList{feature(1)}
    entry_f = b"\x0a\x0a" + b"nnet_input" + b"\x12" + bytes([len(flist_f)]) + flist_f
    feature_i = b"\x1a\x03" + b"\x0a\x01\x03"                           # Feature{int64_list(3){value(1) packed varint 3}}
    flist_i = b"\x0a" + bytes([len(feature_i)]) + feature_i
    entry_i = b"\x0a\x0b" + b"nnet_target" + b"\x12" + bytes([len(flist_i)]) + flist_i
    lists = b"\x0a" + bytes([len(entry_f)]) + entry_f + b"\x0a" + bytes([len(entry_i)]) + entry_i
    want = b"\x12" + bytes([len(lists)]) + lists                        # SequenceExample{feature_lists(2)}
    got = tfr.serialize_sequence_example(np.array([[1.0, 2.0]], np.float32), [3])
    assert got == want
    back = tfr.parse_sequence_example(want)
    assert np.array_equal(back["nnet_input"], [[1.0, 2.0]]) and back["nnet_target"].tolist() == [3]


def test_tfrecord_framing_and_round_trip(tmp_path):
    rng = np.random.RandomState(0)
    x = rng.randn(37, 13).astype(np.float32)
    y = np.array([5, 0, 71, 2 ** 40, -7], dtype=np.int64)
    f = str(tmp_path / "a.tfrecords")
    tfr.write_tfrecord(f, x, y)
    raw = open(f, "rb").read()
    (n,) = struct.unpack("<Q", raw[:8])
    assert len(raw) == n + 16
    mask = lambda c: (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF
    assert struct.unpack("<I", raw[8:12])[0] == mask(crc(raw[:8]))
    assert struct.unpack("<I", raw[12 + n:])[0] == mask(crc(raw[12:12 + n]))
    (r,) = tfr.read_tfrecord(f)
    assert np.array_equal(r["nnet_input"], x) and np.array_equal(r["nnet_target"], y)
    # corruption is detected (tf.errors.DataLossError in the reference's reader)
    bad = bytearray(raw); bad[40] ^= 1
    open(f, "wb").write(bytes(bad))
    with pytest.raises(IOError):
        tfr.read_tfrecord(f)
    # no labels, one frame
    tfr.write_tfrecord(f, x[:1])
    (r,) = tfr.read_tfrecord(f)
    assert np.array_equal(r["nnet_input"], x[:1]) and "nnet_target" not in r


def _splice_loops(x, lc, rc):
    T, D = x.shape
    out = np.zeros((T, D * (1 + lc + rc)), np.float32)
    for t in range(T):
        for c in range(1 + lc + rc):
            out[t, c * D:(c + 1) * D] = x[min(max(t + c - lc, 0), T - 1)]
    return out


@pytest.mark.parametrize("lc,rc,sub", [(0, 0, 0), (1, 1, 3), (2, 0, 2), (0, 3, 0), (1, 1, 1)])
def test_splice_subsample_host_semantics(lc, rc, sub):
    x = np.random.RandomState(1).randn(17, 4).astype(np.float32)
    got = tfr.splice_subsample_host(x, lc, rc, sub)
    want = _splice_loops(x, lc, rc)
    if sub:
        want = want[[i * sub for i in range(want.shape[0] // sub)]]
    assert got.shape == want.shape and np.array_equal(got, want)


def test_dataset_from_tfrecords(tmp_path):
    rng = np.random.RandomState(2)
    utts = [(rng.randn(n, 6).astype(np.float32), rng.randint(0, 9, size=max(1, n // 4))) for n in (9, 14, 5)]
    scp = tmp_path / "feats.scp"
    with open(scp, "w") as fh:
        for i, (x, y) in enumerate(utts):
            p = str(tmp_path / ("u%d.tfrecords" % i))
            tfr.write_tfrecord(p, x, y)
            fh.write("u%d %d %d 1 %s\n" % (i, x.shape[0], x.shape[1], p))
    names, ds, dim = tfr.dataset_from_tfrecords(str(scp), left_context=1, right_context=1, subsample=3)
    assert dim == 18 and len(names) == 3 and ds.device_splice is None
    got = list(ds)
    for (x, y), u in zip(utts, got):
        assert np.array_equal(u["nnet_input"], tfr.splice_subsample_host(x, 1, 1, 3)) and np.array_equal(u["nnet_target"], y)
    # device_splice: raw frames + the recipe for the Session
    _, ds2, dim2 = tfr.dataset_from_tfrecords(str(scp), left_context=1, right_context=1, subsample=3, device_splice=True)
    assert dim2 == 18 and ds2.device_splice == (1, 1, 3) and ds2.raw_input_dim == 6
    assert np.array_equal(next(iter(ds2))["nnet_input"], utts[0][0])
    # shuffle is a seeded permutation of the FILE list (tfrecord.py:87-91)
    n1, _, _ = tfr.dataset_from_tfrecords(str(scp), shuffle=True, seed=7)
    n2, _, _ = tfr.dataset_from_tfrecords(str(scp), shuffle=True, seed=7)
    assert n1 == n2 and sorted(n1) == sorted(names)


def test_kaldi_float_matrix_archive(tmp_path):
    rng = np.random.RandomState(3)
    mats = [("utt_a", rng.randn(4, 3).astype(np.float32)), ("utt-b", rng.randn(1, 5).astype(np.float32))]
    ark, scp = str(tmp_path / "o.ark"), str(tmp_path / "o.scp")
    w = kaldi_io.BaseFloatMatrixWriter("ark,scp:%s,%s" % (ark, scp))
    for k, m in mats:
        assert w.Write(k, m)
    w.Close()
    raw = open(ark, "rb").read()
    head = b"utt_a \0BFM \x04" + struct.pack("<i", 4) + b"\x04" + struct.pack("<i", 3)
    assert raw.startswith(head) and raw[len(head):len(head) + 48] == mats[0][1].tobytes()
    back = kaldi_io.read_float_matrix_ark(ark)
    assert [k for k, _ in back] == ["utt_a", "utt-b"]
    for (_, a), (_, b) in zip(mats, back):
        assert np.array_equal(a, b)
    lines = open(scp).read().split("\n")
    assert lines[0] == "utt_a %s:6" % ark                       # offset of the "\0B" header, right after "utt_a "
    off = int(lines[1].rsplit(":", 1)[1])
    assert raw[off:off + 5] == b"\0BFM "
    with pytest.raises(ValueError):
        kaldi_io.BaseFloatMatrixWriter("ark:%s" % ark).Write("bad key", mats[0][1])
    t = str(tmp_path / "t.ark")
    w = kaldi_io.BaseFloatMatrixWriter("ark,t:%s" % t)
    w.Write("k", np.array([[1.0, 2.5]], np.float32)); w.Close()
    assert open(t).read() == "k  [\n  1.000000 2.500000 ]\n"


def test_native_sequence_example_parser_equals_python_statement():
    """lcb_parse_sequence_example (host C++ in the library, GIL-free) vs the interpreter-level decoder on packed and
    unpacked float encodings, negative / 64-bit labels, empty label lists, and malformed input."""
    rng = np.random.RandomState(3)
    for T, D, L in ((1, 1, 0), (5, 3, 2), (64, 120, 9), (7, 40, 1)):
        x = rng.randn(T, D).astype(np.float32)
        y = rng.randint(-5, 2 ** 40, size=L).astype(np.int64) if L else None
        buf = tfr.serialize_sequence_example(x, y)
        a, b = tfr.parse_sequence_example(buf), tfr.parse_sequence_example_py(buf)
        assert np.array_equal(a["nnet_input"], b["nnet_input"]) and np.array_equal(a["nnet_input"], x)
        assert ("nnet_target" in a) == ("nnet_target" in b)
        if y is not None:
            assert np.array_equal(a["nnet_target"], y)
    # unpacked floats (wire type 5, one value per tag): Feature{float_list{value: 1.5, value: -2.0}}
    fl = b"".join(b"\x0d" + struct.pack("<f", v) for v in (1.5, -2.0))
    feat = b"\x12" + bytes([len(fl)]) + fl
    flist = b"\x0a" + bytes([len(feat)]) + feat
    entry = b"\x0a\x0anet_input"[:0] + b"\x0a" + bytes([10]) + b"nnet_input" + b"\x12" + bytes([len(flist)]) + flist
    fls = b"\x0a" + bytes([len(entry)]) + entry
    msg = b"\x12" + bytes([len(fls)]) + fls
    a, b = tfr.parse_sequence_example(msg), tfr.parse_sequence_example_py(msg)
    assert np.array_equal(a["nnet_input"], np.array([[1.5, -2.0]], np.float32)) and np.array_equal(a["nnet_input"], b["nnet_input"])
    with pytest.raises(ValueError):
        tfr.parse_sequence_example(msg[:-3])                       # truncated
