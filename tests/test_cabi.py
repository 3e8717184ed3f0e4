"""The C-ABI boundary: libnsgpu.so builds for sm_100a, loads, and exports every symbol include/nsgpu.h
declares (no compute calls: this file runs without a GPU)."""
import ctypes
import os
import re

import pytest

from stabilized_navier_stokes_flow_fenicsx_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    _lib.build()
    return _lib.load()


def _declared():
    text = open(os.path.join(ROOT, "include", "nsgpu.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nsgpu_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound(lib):
    names = _declared()
    assert len(names) >= 30
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/nsgpu.h but not exported"
        assert n in _lib.SIGNATURES, f"{n} has no ctypes signature"
    exported = os.popen(f"nm -D --defined-only {_lib.LIB_PATH}").read()
    extra = sorted(set(re.findall(r" T (nsgpu_[a-z0-9_]+)", exported)) - set(names))
    assert not extra, f"exported but undeclared: {extra}"


def test_version_and_loud_failure_without_gpu(lib):
    assert lib.nsgpu_version() >= 100
    ctx = ctypes.c_void_p()
    rc = lib.nsgpu_create(ctypes.byref(ctx), 0)
    if rc == 0:
        assert lib.nsgpu_destroy(ctx) == 0
    else:
        assert rc == -2 and b"no CPU fallback" in lib.nsgpu_last_error(None)
        from stabilized_navier_stokes_flow_fenicsx_b200.assembler import NSAssembler
        import numpy as np
        with pytest.raises(_lib.NsgpuError):
            NSAssembler(np.zeros((4, 3)), np.array([[0, 1, 2, 3]], dtype=np.int32), np.arange(16, dtype=np.int32)[None, :])


def test_binary_is_sm100a_only():
    out = os.popen(f"cuobjdump -lelf {_lib.LIB_PATH} 2>/dev/null").read()
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs
