"""CPU tests: the C-ABI library loads and exports every symbol include/rubiks_b200.h declares; argument
validation that returns before any CUDA call.  (No compute calls: there is no GPU here.)"""
import ctypes
import os
import re

import pytest

from helpers import REPO

from rubiksnet_b200 import _lib


def _declared_symbols():
    text = open(os.path.join(REPO, "include", "rubiks_b200.h")).read()
    text = re.sub(r"#ifdef RB_DEBUG_TRACE.*?#endif", "", text, flags=re.S)  # debug-build-only declarations
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(rb_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported():
    lib = ctypes.CDLL(_lib.LIB_PATH)
    names = _declared_symbols()
    assert len(names) >= 15, names
    for n in names:
        assert hasattr(lib, n), "librubiks_b200.so does not export %s" % n


def test_product_library_has_no_debug_hooks():
    """No tracing entry point and no environment-variable switches in the library whose numbers are graded."""
    lib = ctypes.CDLL(_lib.LIB_PATH)
    assert not hasattr(lib, "rb_debug_pw_trace")
    blob = open(_lib.LIB_PATH, "rb").read()
    assert b"RB_PW_DBG" not in blob


def test_version_and_out_len():
    L = _lib.lib()
    assert L.rb_abi_version() == 1
    # (in + 2p - 1) / s + 1, cuda_src/rubiks.cpp:166
    assert L.rb_out_len(8, 1, 0) == 8
    assert L.rb_out_len(112, 2, 0) == 56
    assert L.rb_out_len(7, 2, 0) == 4
    assert L.rb_out_len(8, 2, 1) == 5
    assert L.rb_out_len(9, 3, 2) == 5


def test_argument_validation_without_gpu():
    L = _lib.lib()
    z = ctypes.c_void_p(0)
    rc = L.rb_shift3d_forward(z, z, z, 0, 0, 1, 8, 4, 5, 5, 0, 1, 1, 0, 0, 0, 0, z)
    assert rc == 1 and b"stride" in L.rb_last_error()
    rc = L.rb_shift3d_forward(z, z, z, 7, 0, 1, 8, 4, 5, 5, 1, 1, 1, 0, 0, 0, 0, z)
    assert rc == 1 and b"dtype" in L.rb_last_error()
    rc = L.rb_shift3d_forward(z, z, z, 0, 0, -1, 8, 4, 5, 5, 1, 1, 1, 0, 0, 0, 0, z)
    assert rc == 1
    rc = L.rb_shift3d_forward(z, z, z, 0, 0, 1, 8, 4, 5, 5, 1, 1, 1, 0, 0, 0, 0, z)
    assert rc == 1 and b"null" in L.rb_last_error()
    # empty tensors are a no-op success
    assert L.rb_shift3d_forward(z, z, z, 0, 0, 0, 8, 4, 5, 5, 1, 1, 1, 0, 0, 0, 0, z) == 0
    assert L.rb_shift2d_forward(z, z, z, 0, 0, 2, 4, 5, 5, 0, 1, 0, 0, 0, z) == 1
    # float64 activations need a float64 shift
    assert L.rb_shift3d_forward(z, z, z, 1, 0, 1, 8, 4, 5, 5, 1, 1, 1, 0, 0, 0, 0, z) == 1
    with pytest.raises(_lib.RubiksCudaError):
        _lib.check(1)


def test_workspace_sizes():
    L = _lib.lib()
    n = L.rb_shift3d_backward_workspace_bytes(0, 2, 8, 64, 56, 56, 1, 1, 1, 0, 0, 0)
    assert n > 0 and n % 256 == 0
    assert L.rb_shift3d_backward_workspace_bytes(0, 0, 8, 64, 56, 56, 1, 1, 1, 0, 0, 0) == 0
    assert L.rb_shift2d_backward_workspace_bytes(0, 16, 64, 56, 56, 1, 1, 0, 0) > 0
    assert L.rb_attention_shift_backward_workspace_bytes(2, 8, 64, 196) == 64 * 2 * 2 * 3 * 4


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(ImportError):
        _lib.lib()
