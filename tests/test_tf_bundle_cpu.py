"""TF checkpoint-V2 tensor bundle (lstm_ctc_b200/tf_bundle.py) -- the format of the reference's `nnet.$iter` models
(bin/nnet-train.py:83-96 `tf.train.Saver(tf.trainable_variables())`).

TensorFlow is not importable here, so the layout is pinned by a HAND-ASSEMBLED index file (every byte written out below from
the published table / BundleEntryProto layout, checksums from an independent bitwise CRC-32C), plus structural properties:
restart points and prefix compression across > 16 keys, multi-block tables, separator keys, corruption detection, and the
reference's variable names / shapes surviving a save -> restore."""
import os
import struct

import numpy as np
import pytest

from lstm_ctc_b200 import tf_bundle as tb


def crc32c_bitwise(data, crc=0):
    """Independent statement of CRC-32C (Castagnoli, reflected polynomial 0x82F63B78)."""
    crc ^= 0xFFFFFFFF
    for b in data:
        crc ^= b
        for _ in range(8):
            crc = (crc >> 1) ^ (0x82F63B78 if crc & 1 else 0)
    return crc ^ 0xFFFFFFFF


def masked(data):
    c = crc32c_bitwise(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def block(contents):
    """contents + trailer: type 0 (uncompressed) + masked CRC-32C over contents and the type byte."""
    return contents + b"\x00" + struct.pack("<I", masked(contents + b"\x00"))


def test_hand_assembled_single_tensor_bundle(tmp_path):
    prefix = str(tmp_path / "nnet.1")
    a = np.array([1.0, 2.0], dtype=np.float32)
    tb.write_bundle(prefix, {"a": a})
    raw = bytes.fromhex("0000803f" "00000040")
    assert open(prefix + ".data-00000-of-00001", "rb").read() == raw

    header = bytes.fromhex("0801" "1a02" "0801")             # num_shards=1 ; version{producer=1}
    entry = (bytes.fromhex("0801")                            # dtype = DT_FLOAT
             + bytes.fromhex("1204" "1202" "0802")            # shape { dim { size: 2 } }
             + bytes.fromhex("2808")                          # size = 8   (offset 0, shard 0: proto3 omits zeros)
             + b"\x35" + struct.pack("<I", masked(raw)))      # crc32c (fixed32, field 6), masked
    data = (bytes([0, 0, len(header)]) + header               # key ""  : shared 0, non-shared 0
            + bytes([0, 1, len(entry)]) + b"a" + entry        # key "a" : shared 0, non-shared 1
            + struct.pack("<II", 0, 1))                       # restart array [0], one restart
    meta = struct.pack("<II", 0, 1)                           # empty metaindex block
    data_handle = bytes([0, len(data)])
    index = bytes([0, 1, len(data_handle)]) + b"b" + data_handle + struct.pack("<II", 0, 1)   # successor("a") = "b"
    meta_off = len(data) + 5
    index_off = meta_off + len(meta) + 5
    footer = bytes([meta_off, len(meta), index_off, len(index)])
    footer += b"\x00" * (40 - len(footer)) + bytes.fromhex("57fb808b247547db")
    want = block(data) + block(meta) + block(index) + footer
    got = open(prefix + ".index", "rb").read()
    assert got == want
    assert open(str(tmp_path / "checkpoint")).read().splitlines()[0] == 'model_checkpoint_path: "nnet.1"'
    back = tb.read_bundle(prefix)
    assert list(back) == ["a"] and back["a"].dtype == np.float32 and np.array_equal(back["a"], a)


def test_crc_matches_independent_implementation():
    rng = np.random.default_rng(0)
    for n in (0, 1, 7, 64, 1000):
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert tb._crc32c(d) == crc32c_bitwise(d)
    m = tb._mask(0x12345678)
    assert tb._unmask(m) == 0x12345678


def test_separator_keys():
    assert tb._shortest_separator(b"abcdef", b"abzz") == b"abd"
    assert tb._shortest_separator(b"abc", b"abd") == b"abc"           # c+1 is not < d
    assert tb._shortest_separator(b"ab", b"abc") == b"ab"             # prefix: unchanged
    assert tb._shortest_separator(b"a\xffb", b"b") == b"a\xffb"
    assert tb._short_successor(b"fd4/frnn4/w_o_diag") == b"g"
    assert tb._short_successor(b"\xff\xffa") == b"\xff\xffb"
    assert tb._short_successor(b"\xff") == b"\xff"


def _reference_variables():
    """Variable names and shapes of a 2-layer peephole BiLSTM + mixture output."""
    from lstm_ctc_b200.blstm import ModelConfig
    from lstm_ctc_b200.model import random_tf_variables
    cfg = ModelConfig({"input_dim": 24, "num_layers": 2, "num_neurons": 64, "num_projects": 64, "num_targets": 12,
                       "use_peepholes": True, "num_experts": 4, "dropout_rate": 1.0})
    return {k: v.numpy() for k, v in random_tf_variables(cfg, 3).items()}


def test_model_variables_round_trip(tmp_path):
    tf_vars = _reference_variables()
    assert "fd0/frnn0/kernel" in tf_vars and "bd1/brnn1/projection/kernel" in tf_vars and "Variable_3" in tf_vars
    assert len(tf_vars) > 16                                          # several restart intervals in the data block
    prefix = str(tmp_path / "exp" / "nnet.init")
    tb.write_bundle(prefix, tf_vars)
    header, entries = tb.read_bundle_index(prefix)
    assert header == {"num_shards": 1, "endianness": 0, "producer": 1, "min_consumer": 0}
    assert list(entries) == sorted(tf_vars, key=lambda s: s.encode())
    off = 0
    for name in entries:                                              # data file: back to back in key order
        e = entries[name]
        assert e["offset"] == off and e["shape"] == tf_vars[name].shape and e["dtype"] == tb.DT_FLOAT and e["shard_id"] == 0
        off += e["size"]
    assert os.path.getsize(tb.data_filename(prefix)) == off
    back = tb.read_bundle(prefix)
    for k, v in tf_vars.items():
        assert back[k].dtype == np.float32 and np.array_equal(back[k], v)
    only = tb.read_bundle(prefix, names=["Variable_1"])
    assert list(only) == ["Variable_1"]
    with pytest.raises(tb.BundleError, match="not found"):
        tb.read_bundle(prefix, names=["fd9/frnn9/kernel"])


def test_prefix_compression_and_restarts():
    keys = [("layer%02d/weights" % i).encode() for i in range(40)]
    items = [(b"", b"h")] + [(k, bytes([i])) for i, k in enumerate(keys)]
    buf = tb.build_table(items)
    assert tb.read_table(buf) == items
    # 41 entries at restart interval 16 -> restarts at entries 0, 16, 32; shared-prefix entries are shorter than their keys
    _, p = tb._read_varint(buf[-48:], 0)
    _, p = tb._read_varint(buf[-48:], p)
    ioff, p = tb._read_varint(buf[-48:], p)
    index_entries = list(tb._block_entries(buf[ioff:ioff + tb._read_varint(buf[-48:], p)[0]]))
    assert len(index_entries) == 1 and index_entries[0][0] == b"m"    # successor("layer39/weights")
    doff, q = tb._read_varint(index_entries[0][1], 0)
    dsize, _ = tb._read_varint(index_entries[0][1], q)
    data = buf[doff:doff + dsize]
    assert struct.unpack_from("<I", data, len(data) - 4)[0] == 3
    assert dsize < sum(len(k) + 4 for k, _ in items)


def test_multi_block_table(monkeypatch):
    monkeypatch.setattr(tb, "BLOCK_SIZE", 64)
    items = [(b"", b"hdr")] + [(("v%03d" % i).encode(), os.urandom(20)) for i in range(50)]
    buf = tb.build_table(items)
    assert tb.read_table(buf) == items
    footer = buf[-48:]
    _, p = tb._read_varint(footer, 0)
    _, p = tb._read_varint(footer, p)
    ioff, p = tb._read_varint(footer, p)
    isize, _ = tb._read_varint(footer, p)
    idx = list(tb._block_entries(buf[ioff:ioff + isize]))
    assert len(idx) > 5
    seps = [k for k, _ in idx]
    assert seps == sorted(seps)
    # each separator is >= every key of its block and < every key of the next one
    pos = 0
    for n, (sep, h) in enumerate(idx):
        off, q = tb._read_varint(h, 0)
        size, _ = tb._read_varint(h, q)
        keys = [k for k, _ in tb._block_entries(buf[off:off + size])]
        assert all(k <= sep for k in keys)
        if n + 1 < len(idx):
            noff, _ = tb._read_varint(idx[n + 1][1], 0)
            nsize = tb._read_varint(idx[n + 1][1], tb._read_varint(idx[n + 1][1], 0)[1])[0]
            assert all(sep < k for k, _ in tb._block_entries(buf[noff:noff + nsize]))
        pos += len(keys)
    assert pos == len(items)


def test_corruption_is_detected(tmp_path):
    prefix = str(tmp_path / "m")
    tb.write_bundle(prefix, {"w": np.arange(12, dtype=np.float32).reshape(3, 4), "s": np.float32(2.5)})
    back = tb.read_bundle(prefix)
    assert back["s"].shape == () and back["s"] == np.float32(2.5) and back["w"].shape == (3, 4)
    idx = bytearray(open(prefix + ".index", "rb").read())
    idx[10] ^= 1
    open(prefix + ".index", "wb").write(bytes(idx))
    with pytest.raises(tb.BundleError, match="checksum"):
        tb.read_bundle(prefix)
    idx[10] ^= 1
    open(prefix + ".index", "wb").write(bytes(idx))
    dat = bytearray(open(tb.data_filename(prefix), "rb").read())
    dat[-1] ^= 0x80
    open(tb.data_filename(prefix), "wb").write(bytes(dat))
    with pytest.raises(tb.BundleError, match="Checksum does not match"):
        tb.read_bundle(prefix)
    assert tb.read_bundle(prefix, verify=False)["w"].shape == (3, 4)
    open(prefix + ".index", "wb").write(bytes(idx[:-1]) + b"\x00")
    with pytest.raises(tb.BundleError, match="magic"):
        tb.read_bundle(prefix)
    with pytest.raises(tb.BundleError, match="no checkpoint"):
        tb.read_bundle(str(tmp_path / "absent"))
    with pytest.raises(tb.BundleError, match="strictly increasing"):
        tb.build_table([(b"b", b""), (b"a", b"")])


def test_other_dtypes_and_offsets(tmp_path):
    prefix = str(tmp_path / "t")
    t = {"step": np.array(7, dtype=np.int64), "h": np.arange(6, dtype=np.float16).reshape(2, 3), "d": np.ones((2, 0, 3), np.float64),
         "i": np.array([-1, 5], dtype=np.int32)}
    tb.write_bundle(prefix, t, update_checkpoint_state=False)
    assert not os.path.exists(str(tmp_path / "checkpoint"))
    back = tb.read_bundle(prefix)
    for k, v in t.items():
        assert back[k].dtype == v.dtype and back[k].shape == v.shape and np.array_equal(back[k], v)


def test_snappy_block_decoding():
    # "abcabcabcabcX": literal "abc", copy(offset 3, length 9) with a 1-byte offset tag, literal "X"
    comp = bytes([13]) + bytes([(3 - 1) << 2]) + b"abc" + bytes([((9 - 4) << 2) | 1 | (0 << 5), 3]) + bytes([0]) + b"X"
    assert tb._snappy_uncompress(comp) == b"abcabcabcabcX"
    # 2-byte-offset copy
    comp = bytes([8]) + bytes([(4 - 1) << 2]) + b"wxyz" + bytes([((4 - 1) << 2) | 2, 4, 0])
    assert tb._snappy_uncompress(comp) == b"wxyzwxyz"
    with pytest.raises(tb.BundleError):
        tb._snappy_uncompress(bytes([5]) + bytes([((4 - 1) << 2) | 2, 9, 0]))
    # a table whose data block is stored snappy-compressed (type 1) reads back
    items = [(b"", b"h"), (b"k", b"vvvvvvvvvvvvvvvv")]
    plain = tb.build_table(items)
    size = plain.index(b"\x00" + struct.pack("<I", masked(plain[:plain.index(struct.pack("<II", 0, 1)) + 8] + b"\x00")))
    blk = plain[:size]
    lit = bytes([len(blk)]) + bytes([(len(blk) - 1) << 2]) + blk          # one literal (< 60 bytes)
    assert len(blk) < 60
    rebuilt = bytearray()
    rebuilt += lit + b"\x01" + struct.pack("<I", masked(lit + b"\x01"))
    meta = struct.pack("<II", 0, 1)
    meta_off = len(rebuilt)
    rebuilt += block(meta)
    h = bytes([0, len(lit)])
    index = bytes([0, 1, len(h)]) + b"l" + h + struct.pack("<II", 0, 1)
    index_off = len(rebuilt)
    rebuilt += block(index)
    footer = bytes([meta_off, len(meta), index_off, len(index)])
    rebuilt += footer + b"\x00" * (40 - len(footer)) + struct.pack("<Q", tb.TABLE_MAGIC)
    assert tb.read_table(bytes(rebuilt)) == items
