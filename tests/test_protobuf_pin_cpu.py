"""Pins the two hand-written protobuf codecs against an INDEPENDENT implementation: Google's protobuf runtime (installed, 6.x)
driven by descriptors built at run time from the published message definitions

  tensorflow/core/example/feature.proto  : BytesList / FloatList / Int64List / Feature / FeatureList / FeatureLists
  tensorflow/core/example/example.proto  : SequenceExample                         (what nnet/tfrecord.py:96-106 parses)
  tensorflow/core/framework/tensor_shape.proto, protobuf/tensor_bundle.proto : TensorShapeProto, BundleHeaderProto,
                                            BundleEntryProto                       (what tf.train.Saver V2 writes, nnet-train.py:83-95)

No TensorFlow artefact exists offline, so this is the strongest pin available for the data formats either side of
the path: bytes produced by our writers parse with protobuf to the same content, and bytes produced by protobuf decode with
our readers (Python statement and the native `lcb_parse_sequence_example`)."""
import numpy as np
import pytest

pb = pytest.importorskip("google.protobuf")
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory  # noqa: E402

from lstm_ctc_b200 import tf_bundle, tfrecord as tfr  # noqa: E402

T = descriptor_pb2.FieldDescriptorProto


def _field(m, name, number, ftype, label=T.LABEL_OPTIONAL, type_name=None, packed=None, oneof=None):
    f = m.field.add()
    f.name, f.number, f.type, f.label = name, number, ftype, label
    if type_name:
        f.type_name = type_name
    if packed is not None:
        f.options.packed = packed
    if oneof is not None:
        f.oneof_index = oneof
    return f


@pytest.fixture(scope="module")
def msgs():
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name, fd.package, fd.syntax = "lcb_tf_formats.proto", "tensorflow", "proto3"
    m = fd.message_type.add(); m.name = "BytesList"; _field(m, "value", 1, T.TYPE_BYTES, T.LABEL_REPEATED)
    m = fd.message_type.add(); m.name = "FloatList"; _field(m, "value", 1, T.TYPE_FLOAT, T.LABEL_REPEATED, packed=True)
    m = fd.message_type.add(); m.name = "Int64List"; _field(m, "value", 1, T.TYPE_INT64, T.LABEL_REPEATED, packed=True)
    m = fd.message_type.add(); m.name = "Feature"
    m.oneof_decl.add().name = "kind"
    _field(m, "bytes_list", 1, T.TYPE_MESSAGE, type_name=".tensorflow.BytesList", oneof=0)
    _field(m, "float_list", 2, T.TYPE_MESSAGE, type_name=".tensorflow.FloatList", oneof=0)
    _field(m, "int64_list", 3, T.TYPE_MESSAGE, type_name=".tensorflow.Int64List", oneof=0)
    m = fd.message_type.add(); m.name = "FeatureList"; _field(m, "feature", 1, T.TYPE_MESSAGE, T.LABEL_REPEATED, ".tensorflow.Feature")
    m = fd.message_type.add(); m.name = "FeatureLists"
    e = m.nested_type.add(); e.name = "FeatureListEntry"; e.options.map_entry = True
    _field(e, "key", 1, T.TYPE_STRING); _field(e, "value", 2, T.TYPE_MESSAGE, type_name=".tensorflow.FeatureList")
    _field(m, "feature_list", 1, T.TYPE_MESSAGE, T.LABEL_REPEATED, ".tensorflow.FeatureLists.FeatureListEntry")
    m = fd.message_type.add(); m.name = "Features"
    e = m.nested_type.add(); e.name = "FeatureEntry"; e.options.map_entry = True
    _field(e, "key", 1, T.TYPE_STRING); _field(e, "value", 2, T.TYPE_MESSAGE, type_name=".tensorflow.Feature")
    _field(m, "feature", 1, T.TYPE_MESSAGE, T.LABEL_REPEATED, ".tensorflow.Features.FeatureEntry")
    m = fd.message_type.add(); m.name = "SequenceExample"
    _field(m, "context", 1, T.TYPE_MESSAGE, type_name=".tensorflow.Features")
    _field(m, "feature_lists", 2, T.TYPE_MESSAGE, type_name=".tensorflow.FeatureLists")
    # tensor_shape.proto / tensor_bundle.proto / versions.proto
    m = fd.message_type.add(); m.name = "TensorShapeProto"
    d = m.nested_type.add(); d.name = "Dim"; _field(d, "size", 1, T.TYPE_INT64); _field(d, "name", 2, T.TYPE_STRING)
    _field(m, "dim", 2, T.TYPE_MESSAGE, T.LABEL_REPEATED, ".tensorflow.TensorShapeProto.Dim"); _field(m, "unknown_rank", 3, T.TYPE_BOOL)
    m = fd.message_type.add(); m.name = "VersionDef"
    _field(m, "producer", 1, T.TYPE_INT32); _field(m, "min_consumer", 2, T.TYPE_INT32)
    _field(m, "bad_consumers", 3, T.TYPE_INT32, T.LABEL_REPEATED, packed=True)
    m = fd.message_type.add(); m.name = "BundleHeaderProto"
    _field(m, "num_shards", 1, T.TYPE_INT32); _field(m, "endianness", 2, T.TYPE_INT32)       # (enum Endianness { LITTLE = 0; BIG = 1 })
    _field(m, "version", 3, T.TYPE_MESSAGE, type_name=".tensorflow.VersionDef")
    m = fd.message_type.add(); m.name = "BundleEntryProto"
    _field(m, "dtype", 1, T.TYPE_INT32)                                                         # (enum DataType; DT_FLOAT = 1)
    _field(m, "shape", 2, T.TYPE_MESSAGE, type_name=".tensorflow.TensorShapeProto")
    _field(m, "shard_id", 3, T.TYPE_INT32); _field(m, "offset", 4, T.TYPE_INT64); _field(m, "size", 5, T.TYPE_INT64)
    _field(m, "crc32c", 6, T.TYPE_FIXED32)
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    get = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName("tensorflow." + n))
    return {n: get(n) for n in ("SequenceExample", "BundleHeaderProto", "BundleEntryProto")}


def test_sequence_example_written_by_us_parses_with_protobuf(msgs):
    rng = np.random.RandomState(1)
    x = rng.randn(9, 5).astype(np.float32)
    y = np.array([3, 0, 70, 2 ** 40, -2], dtype=np.int64)
    se = msgs["SequenceExample"]()
    se.ParseFromString(tfr.serialize_sequence_example(x, y))
    fl = se.feature_lists.feature_list
    assert sorted(fl.keys()) == ["nnet_input", "nnet_target"]
    got = np.array([list(f.float_list.value) for f in fl["nnet_input"].feature], dtype=np.float32)
    assert np.array_equal(got, x)
    lab = [v for f in fl["nnet_target"].feature for v in f.int64_list.value]
    assert lab == y.tolist()


def test_sequence_example_written_by_protobuf_decodes_with_ours(msgs):
    rng = np.random.RandomState(2)
    x = rng.randn(17, 8).astype(np.float32)
    y = [5, 1, 69, -1, 2 ** 33]
    se = msgs["SequenceExample"]()
    se.context.feature["speaker"].bytes_list.value.append(b"ignored by the reference")       # context features are skipped
    for row in x:
        se.feature_lists.feature_list["nnet_input"].feature.add().float_list.value.extend(row.tolist())
    for v in y:                                                       # one Feature per label, like convert-to-tfrecords.py writes them
        se.feature_lists.feature_list["nnet_target"].feature.add().int64_list.value.append(v)
    buf = se.SerializeToString()
    for parse in (tfr.parse_sequence_example, tfr.parse_sequence_example_py):
        r = parse(buf)
        assert np.array_equal(r["nnet_input"], x)
        assert r["nnet_target"].tolist() == y


def test_bundle_protos_match_protobuf(msgs):
    # header: num_shards = 1, LITTLE endian, version.producer = 1 (what tf.train.Saver V2 writes for a single-shard save)
    h = msgs["BundleHeaderProto"]()
    h.num_shards = 1
    h.version.producer = 1
    assert tf_bundle._header_proto() == h.SerializeToString()
    # entries, including a zero-sized dimension, a scalar and large offsets
    for shape, off, size, crc in (((632, 2048), 0, 632 * 2048 * 4, 0x9ABCDEF0), ((2048,), 5177344, 8192, 1), ((), 12, 4, 0xFFFFFFFF),
                                  ((0, 7), 1 << 33, 0, 0)):
        e = msgs["BundleEntryProto"]()
        e.dtype = 1
        for d in shape:
            e.shape.dim.add().size = d
        if not shape:
            e.shape.SetInParent()
        e.offset, e.size, e.crc32c = off, size, crc
        ours = tf_bundle._entry_proto(1, shape, off, size, crc)
        back = msgs["BundleEntryProto"]()
        back.ParseFromString(ours)                                     # our bytes mean the same message ...
        assert back == e
        p = tf_bundle._parse_entry(e.SerializeToString())              # ... and protobuf's bytes decode with our reader
        assert p["dtype"] == 1 and p["shape"] == tuple(shape) and p["offset"] == off and p["size"] == size and p["crc32c"] == crc


def test_snappy_decoder_against_pyarrow():
    """tf_bundle._snappy_uncompress (TF's index blocks may be snappy-compressed) vs an independent encoder: pyarrow's snappy codec."""
    pa = pytest.importorskip("pyarrow")
    if not pa.Codec.is_available("snappy"):
        pytest.skip("pyarrow built without snappy")
    rng = np.random.RandomState(4)
    samples = [b"", b"a", b"fd0/frnn0/kernel" * 200, bytes(rng.randint(0, 4, size=70000).astype(np.uint8)),
               bytes(rng.randint(0, 256, size=5000).astype(np.uint8)), b"\x00" * 100000]
    for raw in samples:
        comp = pa.Codec("snappy").compress(raw, asbytes=True)
        assert tf_bundle._snappy_uncompress(comp) == raw
