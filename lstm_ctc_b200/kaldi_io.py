"""Kaldi float-matrix archive writer (and reader, for tests) -- the output side of bin/nnet-forward.py.

The reference writes posteriors with pyKaldiIO's `BaseFloatMatrixWriter(wspecifier).Write(key, matrix)`
(/root/reference/bin/nnet-forward.py:53-55,93): per utterance `<key> ` then, in binary mode, `\\0B` + `FM ` + int32 rows + int32
cols (each as a size byte 4 followed by the little-endian value) + the row-major float32 payload
(pyKaldiIO/kaldi_matrix.py:280-289, kaldi_table.py:950-975, io_funcs.py InitKaldiOutputStream / WriteBasicType).  That is the
layout Kaldi's `latgen-faster` / `copy-feats` read.  Supported wspecifiers: `ark:<file>`, `ark,scp:<ark>,<scp>`, `ark:-`
(stdout), `ark,t:<file>` (text); options `ark,scp` order-insensitive as in Kaldi."""
import struct
import sys

import numpy as np


def _parse_wspecifier(wspecifier):
    head, _, rest = wspecifier.partition(":")
    opts = [o.strip() for o in head.split(",")]
    if "ark" not in opts and "scp" not in opts:
        raise ValueError("invalid wspecifier: %s" % wspecifier)
    text = "t" in opts
    ark, scp = None, None
    if "ark" in opts and "scp" in opts:
        a, _, b = rest.partition(",")
        ark, scp = (a, b) if opts.index("ark") < opts.index("scp") else (b, a)
    elif "ark" in opts:
        ark = rest
    else:
        raise ValueError("scp-only wspecifiers need one file per key and are not used by nnet-forward: %s" % wspecifier)
    return ark, scp, text


class BaseFloatMatrixWriter:
    def __init__(self, wspecifier):
        self.ark_name, self.scp_name, self.text = _parse_wspecifier(wspecifier)
        self.ark = sys.stdout.buffer if self.ark_name == "-" else open(self.ark_name, "wb")
        self.scp = open(self.scp_name, "w") if self.scp_name else None
        self.pos = 0

    def _w(self, b):
        self.ark.write(b)
        self.pos += len(b)

    def Write(self, key, value):
        if not key or any(c.isspace() for c in key):
            raise ValueError('Using invalid key "%s"' % key)
        m = np.ascontiguousarray(np.asarray(value), dtype="<f4")
        if m.ndim != 2:
            raise ValueError("matrix expected")
        self._w((key + " ").encode())
        if self.scp is not None:
            self.scp.write("%s %s:%d\n" % (key, self.ark_name, self.pos))      # offset of the binary header, as Kaldi writes it
        if self.text:
            if m.shape[0] == 0 or m.shape[1] == 0:
                self._w(b" []\n")
            else:
                rows = ["\n  " + "".join("%f " % v for v in r) for r in m]
                self._w((" [" + "".join(rows) + "]\n").encode())
        else:
            self._w(b"\0B" + b"FM " + b"\x04" + struct.pack("<i", m.shape[0]) + b"\x04" + struct.pack("<i", m.shape[1]))
            self._w(m.tobytes())
        return True

    def Flush(self):
        self.ark.flush()

    def Close(self):
        self.ark.flush()
        if self.ark is not sys.stdout.buffer:
            self.ark.close()
        if self.scp is not None:
            self.scp.close()
        return True


def read_float_matrix_ark(path):
    """Binary `FM ` archive -> list of (key, [rows, cols] float32).  Test helper (the consumer in production is Kaldi)."""
    out = []
    with open(path, "rb") as f:
        data = f.read()
    pos = 0
    while pos < len(data):
        sp = data.index(b" ", pos)
        key = data[pos:sp].decode()
        pos = sp + 1
        if data[pos:pos + 2] != b"\0B" or data[pos + 2:pos + 5] != b"FM ":
            raise IOError("not a binary float-matrix archive at byte %d" % pos)
        pos += 5
        assert data[pos] == 4
        (rows,) = struct.unpack("<i", data[pos + 1:pos + 5])
        assert data[pos + 5] == 4
        (cols,) = struct.unpack("<i", data[pos + 6:pos + 10])
        pos += 10
        m = np.frombuffer(data, dtype="<f4", count=rows * cols, offset=pos).reshape(rows, cols).copy()
        pos += rows * cols * 4
        out.append((key, m))
    return out
