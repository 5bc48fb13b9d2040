"""oracle/ctc.py -- TEST INFRASTRUCTURE.  ctypes wrapper over oracle/ctc_oracle.c (the CPU
restatement of tf.nn.ctc_loss as called at /root/reference/nnet/graph.py:109-114) plus a
pure-NumPy twin for tiny cases and a brute-force path enumerator."""
import ctypes
import itertools
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libctc_oracle.so")
_lib = None


def build_ctc_oracle(force=False):
    """gcc -O2 -fopenmp oracle/ctc_oracle.c -> oracle/_build/libctc_oracle.so"""
    src = os.path.join(_HERE, "ctc_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        os.makedirs(os.path.dirname(_SO), exist_ok=True)
        subprocess.check_call(["gcc", "-O2", "-fopenmp", "-shared", "-fPIC", "-o", _SO, src, "-lm"])
    return _SO


def _load():
    global _lib
    if _lib is None:
        build_ctc_oracle()
        _lib = ctypes.CDLL(_SO)
        for name, fl in (("ctc_oracle_f64", ctypes.c_double), ("ctc_oracle_f32", ctypes.c_float)):
            fn = getattr(_lib, name)
            fn.restype = ctypes.c_int
            fn.argtypes = [ctypes.POINTER(fl), ctypes.POINTER(ctypes.c_int64), ctypes.c_int,
                           ctypes.POINTER(ctypes.c_int32), ctypes.c_int, ctypes.c_int, ctypes.c_int,
                           ctypes.POINTER(fl), ctypes.POINTER(fl), ctypes.c_int]
    return _lib


def ctc_loss_grad(logits, labels, seq_len, dtype=np.float64, nthreads=0):
    """logits [B,T,V]; labels [B,Lmax] int64 (-1 padded); seq_len [B] int32.
    Returns (loss[B], grad[B,T,V]) with TF semantics (blank = V-1)."""
    lib = _load()
    logits = np.ascontiguousarray(logits, dtype=dtype)
    labels = np.ascontiguousarray(labels, dtype=np.int64)
    if labels.ndim == 1:
        labels = labels.reshape(logits.shape[0], -1)
    seq_len = np.ascontiguousarray(seq_len, dtype=np.int32)
    B, T, V = logits.shape
    loss = np.zeros(B, dtype=dtype)
    grad = np.zeros_like(logits)
    fl = ctypes.c_double if dtype == np.float64 else ctypes.c_float
    fn = lib.ctc_oracle_f64 if dtype == np.float64 else lib.ctc_oracle_f32
    rc = fn(logits.ctypes.data_as(ctypes.POINTER(fl)),
            labels.ctypes.data_as(ctypes.POINTER(ctypes.c_int64)), labels.shape[1],
            seq_len.ctypes.data_as(ctypes.POINTER(ctypes.c_int32)), B, T, V,
            loss.ctypes.data_as(ctypes.POINTER(fl)), grad.ctypes.data_as(ctypes.POINTER(fl)),
            int(nthreads))
    if rc != 0:
        raise ValueError("InvalidArgument: label out of range [0, num_classes-1)")
    return loss, grad


def ctc_loss_numpy(logits, label, T_b):
    """Pure-NumPy single-utterance twin (probability space, fp64) used to cross-check the C
    file on tiny inputs.  logits [T,V]; label: list[int]."""
    x = np.asarray(logits, dtype=np.float64)[:T_b]
    V = x.shape[1]
    blank = V - 1
    y = np.exp(x - x.max(1, keepdims=True))
    y /= y.sum(1, keepdims=True)
    ext = [blank]
    for l in label:
        ext += [int(l), blank]
    S = len(ext)
    al = np.zeros((T_b, S))
    al[0, 0] = y[0, blank]
    if S > 1:
        al[0, 1] = y[0, ext[1]]
    for t in range(1, T_b):
        for s in range(S):
            a = al[t - 1, s]
            if s >= 1:
                a += al[t - 1, s - 1]
            if s >= 2 and ext[s] != blank and ext[s] != ext[s - 2]:
                a += al[t - 1, s - 2]
            al[t, s] = a * y[t, ext[s]]
    p = al[T_b - 1, S - 1] + (al[T_b - 1, S - 2] if S > 1 else 0.0)
    return -np.log(p) if p > 0 else np.inf


def ctc_brute_force(logits, label, T_b):
    """Sum of path probabilities over all V^T alignments that collapse to `label`."""
    x = np.asarray(logits, dtype=np.float64)[:T_b]
    V = x.shape[1]
    blank = V - 1
    y = np.exp(x - x.max(1, keepdims=True))
    y /= y.sum(1, keepdims=True)
    tot = 0.0
    for path in itertools.product(range(V), repeat=T_b):
        out, prev = [], None
        for c in path:
            if c != prev and c != blank:
                out.append(c)
            prev = c
        if out == list(label):
            pr = 1.0
            for t, c in enumerate(path):
                pr *= y[t, c]
            tot += pr
    return -np.log(tot) if tot > 0 else np.inf
