"""ctypes binding of liblstm_ctc_b200.so (the C ABI declared in include/lstm_ctc_b200.h).

There is NO fallback: if the shared library is missing or a symbol is absent this module raises.
The library is built in-tree by `build()` (nvcc, sm_100a only) so that it travels with the repo
snapshot to the GPU box.
"""
import ctypes
import os
import re
import subprocess

_PKG = os.path.dirname(os.path.abspath(__file__))
_ROOT = os.path.dirname(_PKG)
_CSRC = os.path.join(_PKG, "csrc")
_SO = os.path.join(_PKG, "liblstm_ctc_b200.so")
_HEADER = os.path.join(_ROOT, "include", "lstm_ctc_b200.h")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
              "--shared", "-Xcompiler", "-fPIC"]


def _sources():
    return [os.path.join(_CSRC, f) for f in sorted(os.listdir(_CSRC)) if f.endswith((".cu", ".cuh", ".h"))] + [_HEADER]


def build(force=False, verbose=False):
    """nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo ... csrc/lcb_all.cu -> liblstm_ctc_b200.so"""
    if not force and os.path.exists(_SO):
        newest = max(os.path.getmtime(s) for s in _sources())
        if os.path.getmtime(_SO) >= newest:
            return _SO
    cmd = ["nvcc"] + NVCC_FLAGS + ["-I", os.path.join(_ROOT, "include"), "-I", _CSRC,
                                   "-o", _SO, os.path.join(_CSRC, "lcb_all.cu")]
    if verbose:
        cmd.insert(1, "-Xptxas=-v")
        print(" ".join(cmd))
    subprocess.check_call(cmd)
    return _SO


def declared_symbols():
    """Every function name include/lstm_ctc_b200.h declares."""
    txt = open(_HEADER).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(lcb_[a-z0-9_]+)\s*\(", txt)))


_lib = None

c_int = ctypes.c_int
c_void_p = ctypes.c_void_p
c_size_t = ctypes.c_size_t
c_float = ctypes.c_float

_SIGS = {
    "lcb_version": (c_int, []),
    "lcb_status_string": (ctypes.c_char_p, [c_int]),
    "lcb_device_error": (c_int, [c_int]),
    "lcb_launch_count": (ctypes.c_longlong, [c_int]),
    "lcb_launch_count_add": (None, [ctypes.c_longlong]),
    "lcb_ctc_workspace_bytes": (c_size_t, [c_int] * 4),
    "lcb_ctc_loss_grad_f32": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                                      c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "lcb_ctc_loss_grad_f32_layout": (c_int, [c_void_p, c_void_p, c_int, c_void_p, c_int, c_int, c_int,
                                             c_void_p, c_void_p, c_void_p, c_size_t, c_int, c_void_p]),
    "lcb_ctc_status": (c_int, [c_void_p, c_void_p]),
    "lcb_gemm_bf16": (c_int, [c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_int,
                              c_void_p, c_int, c_int, c_void_p, c_int, c_void_p]),
    "lcb_gemm_bf16_simt_check": (c_int, [c_int, c_int, c_int, c_void_p, c_int, c_int, c_void_p, c_int, c_int,
                                         c_void_p, c_int, c_int, c_void_p, c_int, c_void_p]),
    "lcb_lstm_rec_config": (c_int, [c_int, ctypes.POINTER(c_int), ctypes.POINTER(c_int)]),
    "lcb_lstm_rec_max_clusters": (c_int, [c_int, c_int]),
    "lcb_debug_rec_profile": (c_int, [c_void_p, c_int]),
    "lcb_debug_fwd_layout": (c_int, [c_int]),
    "lcb_debug_ctc_min_spt": (c_int, [c_int]),
    "lcb_lstm_rec_workspace_bytes": (c_size_t, [c_int, c_int]),
    "lcb_lstm_rec_grid": (c_int, [c_int, c_int, c_int, c_int]),
    "lcb_lstm_rec_fwd": (c_int, [c_void_p] * 9 + [c_int, c_int, c_int, c_int, c_float, c_void_p, c_size_t, c_void_p]),
    "lcb_lstm_rec_fwd_range": (c_int, [c_void_p] * 9 + [c_int, c_int, c_int, c_int, c_float, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "lcb_store_i32": (c_int, [c_void_p, c_int, c_void_p]),
    "lcb_lstm_rec_fwd_range_hl": (c_int, [c_void_p] * 11 + [c_int, c_int, c_int, c_int, c_float, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "lcb_lstm_rec_fwd_range_pg": (c_int, [c_void_p, c_int] + [c_void_p] * 11 + [c_int, c_int, c_int, c_int, c_float, c_int, c_int, c_void_p,
                                          c_void_p, c_size_t, c_void_p]),
    "lcb_lstm_rec_fwd_progress_words": (c_int, [c_int, c_int, c_int]),
    "lcb_lstm_rec_bwd": (c_int, [c_void_p] * 9 + [c_int, c_int, c_int, c_int, c_void_p, c_size_t, c_void_p]),
    "lcb_lstm_rec_bwd_range": (c_int, [c_void_p] * 9 + [c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_size_t, c_void_p]),
    "lcb_lstm_rec_bwd_range_pg": (c_int, [c_void_p] * 9 + [c_int, c_int, c_int, c_int, c_int, c_int, c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "lcb_lstm_rec_bwd_progress_words": (c_int, [c_int, c_int, c_int]),
    "lcb_wait_progress": (c_int, [c_void_p, c_int, c_int, c_void_p]),
    "lcb_lstm_rec_bwd_can_split": (c_int, [c_int]),
    "lcb_pack_input": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "lcb_cast_f32_16": (c_int, [c_void_p, c_void_p, c_int, c_size_t, c_void_p]),
    "lcb_device_sm_count": (c_int, []),
    "lcb_gemm16": (c_int, [c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int,
                           c_void_p, c_int, c_int, c_void_p, c_int, c_int, c_void_p]),
    "lcb_gemm16_dropout": (c_int, [c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int,
                                   c_void_p, c_int, c_int, c_void_p, c_int, c_float, ctypes.c_ulonglong, ctypes.c_ulonglong, c_int, c_void_p]),
    "lcb_gemm16_twin": (c_int, [c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int,
                                c_void_p, c_int, c_int, c_void_p, c_void_p, c_int, c_float, ctypes.c_ulonglong, ctypes.c_ulonglong, c_int, c_void_p]),
    "lcb_gemm16_simt_check": (c_int, [c_int, c_int, c_int, c_void_p, c_int, c_int, c_int, c_void_p, c_int, c_int, c_int,
                                      c_void_p, c_int, c_int, c_void_p, c_int, c_void_p]),
    "lcb_split_f32_bf16": (c_int, [c_void_p, c_void_p, c_void_p, c_size_t, c_void_p]),
    "lcb_f16_to_bf16": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "lcb_output_fwd": (c_int, [c_void_p, c_int, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_float,
                               c_float, ctypes.c_ulonglong, c_void_p]),
    "lcb_mos_bwd_dz": (c_int, [c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float,
                               c_float, ctypes.c_ulonglong, c_void_p]),
    "lcb_dropout16": (c_int, [c_void_p, c_int, c_size_t, c_float, ctypes.c_ulonglong, c_void_p]),
    "lcb_dropout_mask": (c_int, [c_void_p, c_size_t, c_float, ctypes.c_ulonglong, c_void_p]),
    "lcb_pack_dlogits": (c_int, [c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_void_p]),
    "lcb_optimizer_step": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, ctypes.c_longlong, c_int, c_float, ctypes.c_longlong,
                                   c_float, c_float, c_float, c_float, c_float, c_float, ctypes.POINTER(ctypes.c_longlong), c_int,
                                   c_void_p, c_void_p, c_void_p]),
    "lcb_greedy_decode": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_void_p]),
    "lcb_posterior": (c_int, [c_void_p, c_void_p, ctypes.c_longlong, c_int, c_float, c_int, c_void_p, c_int, c_void_p]),
    "lcb_add_f16": (c_int, [c_void_p, c_void_p, c_size_t, c_void_p]),
    "lcb_masked_add16": (c_int, [c_void_p, c_int, c_void_p, c_int, ctypes.c_longlong, c_int, c_int, c_float, ctypes.c_ulonglong, ctypes.c_ulonglong, c_int, c_void_p]),
    "lcb_label_smooth": (c_int, [c_void_p, c_void_p, ctypes.c_longlong, c_int, c_float, c_void_p, c_void_p, c_void_p]),
    "lcb_splice_subsample": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_void_p]),
    "lcb_crc32c": (ctypes.c_uint32, [c_void_p, c_size_t, ctypes.c_uint32]),
    "lcb_parse_sequence_example": (c_int, [c_void_p, c_size_t, c_void_p, c_size_t, c_void_p, c_size_t,
                                           ctypes.POINTER(ctypes.c_longlong), ctypes.POINTER(ctypes.c_longlong),
                                           ctypes.POINTER(ctypes.c_longlong)]),
    "lcb_colsum": (c_int, [c_void_p, c_int, c_int, c_int, c_int, c_void_p, c_void_p]),
}


def register_sigs(extra):
    _SIGS.update(extra)


def lib():
    """Load the shared library (building it if sources are newer and nvcc is on PATH)."""
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            try:
                build()
            except Exception as e:  # no nvcc on this box and no prebuilt .so: fail loudly
                raise RuntimeError(
                    "liblstm_ctc_b200.so is missing and could not be built (%s). "
                    "There is no CPU fallback: run __graft_entry__.build() first." % (e,))
        L = ctypes.CDLL(_SO)
        for name, (res, args) in _SIGS.items():
            fn = getattr(L, name)          # AttributeError if the .so lacks a declared symbol
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


class LcbError(RuntimeError):
    def __init__(self, status, where=""):
        self.status = status
        msg = lib().lcb_status_string(status).decode()
        super().__init__("%s: %s (status %d)" % (where, msg, status))


class InvalidArgumentError(LcbError, ValueError):
    """Mirrors tf.errors.InvalidArgumentError raised by tf.nn.ctc_loss for bad labels."""


def check(status, where=""):
    if status != 0:
        if status == -6:
            raise InvalidArgumentError(status, where)
        raise LcbError(status, where)


def ptr(t):
    """data pointer of a CUDA tensor (or None)."""
    if t is None:
        return None
    assert t.is_cuda, "lstm_ctc_b200 kernels take device tensors only (no CPU fallback)"
    return ctypes.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
