"""CPU: the C-ABI shared library builds for sm_100a, loads, and exports every symbol that
include/lstm_ctc_b200.h declares (no compute calls here -- there is no GPU)."""
import ctypes
import os
import subprocess

from lstm_ctc_b200 import _lib


def test_library_builds_and_exports_declared_symbols():
    so = _lib.build()
    assert os.path.exists(so)
    L = ctypes.CDLL(so)
    declared = _lib.declared_symbols()
    assert len(declared) >= 8
    for name in declared:
        assert hasattr(L, name), "missing export: " + name
    # and every symbol the Python side binds is declared in the header
    for name in _lib._SIGS:
        assert name in declared, "bound but undeclared: " + name


def test_status_strings_and_version():
    L = _lib.lib()
    assert L.lcb_version() >= 100
    assert L.lcb_status_string(0) == b"ok"
    assert b"InvalidArgument" in L.lcb_status_string(-6)


def test_argument_validation_without_gpu():
    L = _lib.lib()
    assert L.lcb_ctc_workspace_bytes(0, 10, 5, 3) == 0
    assert L.lcb_ctc_workspace_bytes(4, 100, 30, 10) > 0
    assert L.lcb_ctc_loss_grad_f32(None, None, 0, None, 1, 1, 2, None, None, None, 0, None) == -1
    assert L.lcb_gemm_bf16(0, 1, 1, None, 8, 0, None, 8, 0, None, 8, 0, None, 0, None) == -1


def test_sass_has_blackwell_mnemonics():
    out = subprocess.run(["cuobjdump", "-sass", _lib._SO], capture_output=True, text=True).stdout
    for m in ("UTCHMMA", "UTMALDG", "LDTM"):
        assert m in out, m
