"""TF-free reader / writer of TensorFlow's checkpoint-V2 "tensor bundle" -- the on-disk format of the reference's models.

The reference saves and restores with `tf.train.Saver(tf.trainable_variables())` (bin/nnet-init.py:77-79,
bin/nnet-train.py:83-96, bin/nnet-validate.py, bin/nnet-forward.py:60-62): `saver.save(sess, "exp/nnet.3")` leaves

    exp/nnet.3.index                   key -> BundleEntryProto table (an SSTable in TF's leveldb-derived "table" format)
    exp/nnet.3.data-00000-of-00001     the tensors' raw little-endian bytes, back to back in key order
    exp/nnet.3.meta                    the MetaGraphDef (graph structure; never read back by the reference, not written here)
    exp/checkpoint                     CheckpointState text proto naming the latest prefix

and the shell drivers only ever pass the PREFIX around (scripts/train.sh:123-124,164,230).  This module writes and reads
the first two (and `checkpoint`), so a model trained by either side can be loaded by the other.  TensorFlow is not
installable in this image, so the layout is restated from the published format of TF r1.8
(tensorflow/core/util/tensor_bundle/tensor_bundle.{h,cc}, tensorflow/core/protobuf/tensor_bundle.proto,
tensorflow/core/lib/io/{format,block_builder,table_builder}.cc) and pinned by hand-assembled byte strings in
tests/test_tf_bundle_cpu.py -- NOT by a TF binary ("parity unpinned" at that boundary).

Table format (as leveldb's, with CRC-32C block trailers):
    data block*   entries `varint shared | varint non_shared | varint value_len | key suffix | value`, prefix-compressed
                  against the previous key, restart (shared = 0) every 16 entries; then u32 restart offsets, u32 count
    metaindex     an empty block
    index block   one entry per data block: separator key -> BlockHandle(varint offset, varint size); restart interval 1
    every block is followed by a 5-byte trailer: compression type (0 none / 1 snappy), masked CRC-32C of block + type
    footer        metaindex handle, index handle, zero padding to 40 bytes, magic 0xdb4775248b80fb57 (little-endian)
Keys: "" -> BundleHeaderProto{num_shards=1, endianness=LITTLE, version{producer=1}}; variable name -> BundleEntryProto
{dtype, shape, shard_id, offset, size, crc32c (masked CRC-32C of the tensor bytes)}.
"""
import os
import struct

import numpy as np

from .tfrecord import _crc32c, _ld, _read_varint, _varint

TABLE_MAGIC = 0xDB4775248B80FB57
BLOCK_RESTART_INTERVAL = 16          # table::Options::block_restart_interval
BLOCK_SIZE = 262144                  # table::Options::block_size
_MASK_DELTA = 0xA282EAD8

# tensorflow/core/framework/types.proto
DT_FLOAT, DT_DOUBLE, DT_INT32, DT_INT64, DT_BFLOAT16, DT_HALF = 1, 2, 3, 9, 14, 19
_NP_OF_DT = {DT_FLOAT: np.dtype("<f4"), DT_DOUBLE: np.dtype("<f8"), DT_INT32: np.dtype("<i4"), DT_INT64: np.dtype("<i8"),
             DT_HALF: np.dtype("<f2")}
_DT_OF_NP = {v: k for k, v in _NP_OF_DT.items()}


class BundleError(ValueError):
    """Malformed or unsupported bundle (the analogue of TF's DataLoss / InvalidArgument / NotFound statuses)."""


def _mask(crc):
    return (((crc >> 15) | (crc << 17)) + _MASK_DELTA) & 0xFFFFFFFF


def _unmask(m):
    rot = (m - _MASK_DELTA) & 0xFFFFFFFF
    return ((rot >> 17) | (rot << 15)) & 0xFFFFFFFF


# --------------------------------------------------------------------------------------------------------- table writer
class _BlockBuilder:
    def __init__(self, restart_interval):
        self.interval = restart_interval
        self.buf = bytearray()
        self.restarts = [0]
        self.counter = 0
        self.last_key = b""

    def add(self, key, value):
        shared = 0
        if self.counter < self.interval:
            n = min(len(self.last_key), len(key))
            while shared < n and self.last_key[shared] == key[shared]:
                shared += 1
        else:
            self.restarts.append(len(self.buf))
            self.counter = 0
        self.buf += _varint(shared) + _varint(len(key) - shared) + _varint(len(value)) + key[shared:] + value
        self.last_key = key
        self.counter += 1

    def size_estimate(self):
        return len(self.buf) + 4 * len(self.restarts) + 4

    def empty(self):
        return not self.buf

    def finish(self):
        return bytes(self.buf) + b"".join(struct.pack("<I", r) for r in self.restarts) + struct.pack("<I", len(self.restarts))


def _shortest_separator(start, limit):
    """table_builder.cc FindShortestSeparator: a short key k with start <= k < limit."""
    n = min(len(start), len(limit))
    d = 0
    while d < n and start[d] == limit[d]:
        d += 1
    if d < n and start[d] < 0xFF and start[d] + 1 < limit[d]:
        return start[:d] + bytes([start[d] + 1])
    return start


def _short_successor(key):
    """table_builder.cc FindShortSuccessor: a short key k >= key."""
    for i, b in enumerate(key):
        if b != 0xFF:
            return key[:i] + bytes([b + 1])
    return key


def _handle(offset, size):
    return _varint(offset) + _varint(size)


def build_table(items):
    """Serialise sorted (key bytes, value bytes) pairs as an uncompressed table file (BundleWriter::Finish uses
    table::kNoCompression).  Returns the file's bytes."""
    out = bytearray()
    index = _BlockBuilder(1)
    data = _BlockBuilder(BLOCK_RESTART_INTERVAL)
    pending = None                     # (last key of the flushed block, its handle): its index entry waits for the next key
    last_key = None

    def write_block(contents):
        off = len(out)
        crc = _crc32c(contents + b"\x00")
        out.extend(contents)
        out.extend(b"\x00" + struct.pack("<I", _mask(crc)))
        return _handle(off, len(contents))

    for key, value in items:
        if last_key is not None and not key > last_key:
            raise BundleError("table keys must be strictly increasing: %r after %r" % (key, last_key))
        if pending is not None:
            index.add(_shortest_separator(pending[0], key), pending[1])
            pending = None
        data.add(key, value)
        last_key = key
        if data.size_estimate() >= BLOCK_SIZE:
            pending = (last_key, write_block(data.finish()))
            data = _BlockBuilder(BLOCK_RESTART_INTERVAL)
    if not data.empty():
        pending = (last_key, write_block(data.finish()))
    meta_handle = write_block(_BlockBuilder(BLOCK_RESTART_INTERVAL).finish())
    if pending is not None:
        index.add(_short_successor(pending[0]), pending[1])
    index_handle = write_block(index.finish())
    footer = meta_handle + index_handle
    footer += b"\x00" * (40 - len(footer)) + struct.pack("<Q", TABLE_MAGIC)
    out.extend(footer)
    return bytes(out)


# --------------------------------------------------------------------------------------------------------- table reader
def _snappy_uncompress(src):
    """Raw snappy block format (TF's default table compression; BundleWriter does not use it, MergeBundles-era tools may)."""
    n, pos = _read_varint(src, 0)
    out = bytearray()
    while pos < len(src):
        tag = src[pos]
        pos += 1
        kind = tag & 3
        if kind == 0:
            ln = tag >> 2
            if ln >= 60:
                nb = ln - 59
                ln = int.from_bytes(src[pos:pos + nb], "little")
                pos += nb
            ln += 1
            out += src[pos:pos + ln]
            pos += ln
            continue
        if kind == 1:
            ln = ((tag >> 2) & 7) + 4
            off = ((tag >> 5) << 8) | src[pos]
            pos += 1
        elif kind == 2:
            ln = (tag >> 2) + 1
            off = src[pos] | (src[pos + 1] << 8)
            pos += 2
        else:
            ln = (tag >> 2) + 1
            off = int.from_bytes(src[pos:pos + 4], "little")
            pos += 4
        if off == 0 or off > len(out):
            raise BundleError("corrupt snappy block")
        for _ in range(ln):                   # byte-wise: copies may overlap their own output
            out.append(out[-off])
    if len(out) != n:
        raise BundleError("corrupt snappy block: %d bytes, header says %d" % (len(out), n))
    return bytes(out)


def _read_block(buf, offset, size, verify):
    if offset + size + 5 > len(buf):
        raise BundleError("truncated table: block [%d, +%d) past end of file" % (offset, size))
    contents = buf[offset:offset + size]
    ctype = buf[offset + size]
    if verify:
        want = struct.unpack_from("<I", buf, offset + size + 1)[0]
        if _unmask(want) != _crc32c(buf[offset:offset + size + 1]):
            raise BundleError("block checksum mismatch at offset %d" % offset)
    if ctype == 1:
        contents = _snappy_uncompress(contents)
    elif ctype != 0:
        raise BundleError("bad block type %d" % ctype)
    return contents


def _block_entries(block):
    if len(block) < 4:
        raise BundleError("bad block contents")
    nrestart = struct.unpack_from("<I", block, len(block) - 4)[0]
    end = len(block) - 4 - 4 * nrestart
    if end < 0:
        raise BundleError("bad block contents")
    pos, key = 0, b""
    while pos < end:
        shared, pos = _read_varint(block, pos)
        non_shared, pos = _read_varint(block, pos)
        vlen, pos = _read_varint(block, pos)
        if shared > len(key) or pos + non_shared + vlen > end:
            raise BundleError("corrupt block entry")
        key = key[:shared] + block[pos:pos + non_shared]
        pos += non_shared
        yield key, block[pos:pos + vlen]
        pos += vlen


def read_table(buf, verify=True):
    """All (key, value) pairs of a table file, in key order."""
    if len(buf) < 48:
        raise BundleError("file is too short to be a table")
    footer = buf[-48:]
    if struct.unpack_from("<Q", footer, 40)[0] != TABLE_MAGIC:
        raise BundleError("not a table (bad magic number)")
    pos = 0
    _, pos = _read_varint(footer, pos)
    _, pos = _read_varint(footer, pos)
    ioff, pos = _read_varint(footer, pos)
    isize, pos = _read_varint(footer, pos)
    items = []
    for _, h in _block_entries(_read_block(buf, ioff, isize, verify)):
        off, p = _read_varint(h, 0)
        size, p = _read_varint(h, p)
        items.extend(_block_entries(_read_block(buf, off, size, verify)))
    return items


# ------------------------------------------------------------------------------------------------------------- protobuf
def _pb_varint_field(field, v):
    return _varint(field << 3) + _varint(v)


def _header_proto():
    # num_shards = 1 ; endianness = LITTLE (0: omitted) ; version { producer = 1 (kTensorBundleVersion) ; min_consumer = 0 }
    return _pb_varint_field(1, 1) + _ld(3, _pb_varint_field(1, 1))


def _entry_proto(dtype, shape, offset, size, crc_masked, shard_id=0):
    out = _pb_varint_field(1, dtype)
    out += _ld(2, b"".join(_ld(2, _pb_varint_field(1, int(d)) if d else b"") for d in shape))   # proto3: a zero dim size is omitted
    if shard_id:
        out += _pb_varint_field(3, shard_id)
    if offset:
        out += _pb_varint_field(4, offset)
    if size:
        out += _pb_varint_field(5, size)
    if crc_masked:
        out += _varint((6 << 3) | 5) + struct.pack("<I", crc_masked)
    return out


def _parse_fields(buf):
    """{field: [value]} of one message; varint -> int, fixed32 -> int, length-delimited -> bytes."""
    pos, out = 0, {}
    while pos < len(buf):
        tag, pos = _read_varint(buf, pos)
        field, wt = tag >> 3, tag & 7
        if wt == 0:
            v, pos = _read_varint(buf, pos)
        elif wt == 2:
            n, pos = _read_varint(buf, pos)
            v = buf[pos:pos + n]
            pos += n
        elif wt == 5:
            v = struct.unpack_from("<I", buf, pos)[0]
            pos += 4
        elif wt == 1:
            v = struct.unpack_from("<Q", buf, pos)[0]
            pos += 8
        else:
            raise BundleError("unsupported protobuf wire type %d" % wt)
        out.setdefault(field, []).append(v)
    return out


def _parse_entry(buf):
    f = _parse_fields(buf)
    shape = []
    for sh in f.get(2, []):
        sf = _parse_fields(sh)
        if sf.get(3, [0])[-1]:
            raise BundleError("tensor of unknown rank")
        for dim in sf.get(2, []):
            d = _parse_fields(dim).get(1, [0])[-1]
            shape.append(d - (1 << 64) if d >> 63 else d)
    return {"dtype": f.get(1, [0])[-1], "shape": tuple(shape), "shard_id": f.get(3, [0])[-1], "offset": f.get(4, [0])[-1],
            "size": f.get(5, [0])[-1], "crc32c": f.get(6, [0])[-1], "sliced": bool(f.get(7))}


# --------------------------------------------------------------------------------------------------------------- bundle
def data_filename(prefix, shard=0, num_shards=1):
    return "%s.data-%05d-of-%05d" % (prefix, shard, num_shards)


def index_filename(prefix):
    return prefix + ".index"


def bundle_exists(prefix):
    return os.path.exists(index_filename(prefix))


def write_bundle(prefix, tensors, update_checkpoint_state=True):
    """Write {name: array} as `<prefix>.index` + `<prefix>.data-00000-of-00001`, names in byte order (what
    BaseSaverBuilder's name-sorted saveables and BundleWriter's std::map produce), then -- as Saver.save does -- point
    `<dir>/checkpoint` at it.  Arrays may be numpy arrays or torch tensors of a supported dtype."""
    named = []
    for name, t in tensors.items():
        a = t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)
        if a.dtype.newbyteorder("<") not in _DT_OF_NP:
            raise BundleError("unsupported dtype %s for %r" % (a.dtype, name))
        key = name.encode("utf-8")
        if not key:
            raise BundleError("empty tensor name")
        named.append((key, a.astype(a.dtype.newbyteorder("<"), order="C", copy=False)))
    named.sort(key=lambda kv: kv[0])
    d = os.path.dirname(prefix)
    if d:
        os.makedirs(d, exist_ok=True)
    items = [(b"", _header_proto())]
    offset = 0
    tmp = data_filename(prefix) + ".tempstate"
    with open(tmp, "wb") as f:
        for key, a in named:
            raw = a.tobytes()
            f.write(raw)
            items.append((key, _entry_proto(_DT_OF_NP[a.dtype], a.shape, offset, len(raw), _mask(_crc32c(raw)))))
            offset += len(raw)
    os.replace(tmp, data_filename(prefix))
    tmp = index_filename(prefix) + ".tempstate"
    with open(tmp, "wb") as f:
        f.write(build_table(items))
    os.replace(tmp, index_filename(prefix))
    if update_checkpoint_state:
        base = os.path.basename(prefix)
        with open(os.path.join(d or ".", "checkpoint"), "w") as f:
            f.write('model_checkpoint_path: "%s"\nall_model_checkpoint_paths: "%s"\n' % (base, base))
    return prefix


def read_bundle_index(prefix, verify=True):
    """(header fields, {name: entry dict}) of `<prefix>.index`."""
    path = index_filename(prefix)
    if not os.path.exists(path):
        raise BundleError("no checkpoint at %r (missing %s)" % (prefix, path))
    with open(path, "rb") as f:
        items = read_table(f.read(), verify)
    if not items or items[0][0] != b"":
        raise BundleError("bundle has no header entry")
    hf = _parse_fields(items[0][1])
    header = {"num_shards": hf.get(1, [0])[-1], "endianness": hf.get(2, [0])[-1],
              "producer": _parse_fields(hf[3][-1]).get(1, [0])[-1] if 3 in hf else 0,
              "min_consumer": _parse_fields(hf[3][-1]).get(2, [0])[-1] if 3 in hf else 0}
    if header["endianness"] != 0:
        raise BundleError("big-endian bundles are not supported")
    if header["min_consumer"] > 1:
        raise BundleError("bundle needs a consumer of version >= %d" % header["min_consumer"])
    return header, {k.decode("utf-8"): _parse_entry(v) for k, v in items[1:]}


def read_bundle(prefix, names=None, verify=True):
    """{name: numpy array} of a bundle (all tensors, or `names`); checksums verified like BundleReader::GetValue."""
    header, entries = read_bundle_index(prefix, verify)
    if names is None:
        names = list(entries)
    out, files = {}, {}
    try:
        for name in names:
            if name not in entries:
                raise BundleError("Key %s not found in checkpoint %r" % (name, prefix))
            e = entries[name]
            if e["sliced"]:
                raise BundleError("partitioned variable %r is not supported" % name)
            if e["dtype"] not in _NP_OF_DT:
                raise BundleError("unsupported dtype enum %d for %r" % (e["dtype"], name))
            dt = _NP_OF_DT[e["dtype"]]
            count = int(np.prod(e["shape"], dtype=np.int64)) if e["shape"] else 1
            if count * dt.itemsize != e["size"]:
                raise BundleError("%r: shape %s needs %d bytes, entry says %d" % (name, e["shape"], count * dt.itemsize, e["size"]))
            sid = e["shard_id"]
            if sid not in files:
                files[sid] = open(data_filename(prefix, sid, max(1, header["num_shards"])), "rb")
            f = files[sid]
            f.seek(e["offset"])
            raw = f.read(e["size"])
            if len(raw) != e["size"]:
                raise BundleError("%r: data file is truncated" % name)
            if verify and _unmask(e["crc32c"]) != _crc32c(raw):
                raise BundleError("Checksum does not match for tensor %r" % name)
            out[name] = np.frombuffer(raw, dtype=dt).reshape(e["shape"]).copy()
    finally:
        for f in files.values():
            f.close()
    return out
