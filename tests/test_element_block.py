"""The factorised entity blocks the row-owner kernel evaluates (csrc/element_block.cuh) against the direct-quadrature rows of
csrc/element_generic.cuh (the formulation the oracle parity tests pin), compiled with g++: every element pair, every form."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def host():
    out = os.path.join(ROOT, "tests", "host", "libblock_host.so")
    src = os.path.join(ROOT, "tests", "host", "block_host.cpp")
    csrc = os.path.join(ROOT, "stabilized_navier_stokes_flow_fenicsx_b200", "csrc")
    newest = max(os.path.getmtime(p) for p in [src] + [os.path.join(csrc, h) for h in ("element_block.cuh", "element_shared.cuh", "element_generic.cuh")])
    if not os.path.exists(out) or os.path.getmtime(out) < newest:
        subprocess.check_call(["g++", "-O1", "-shared", "-fPIC", "-x", "c++", src, "-o", out])
    return ctypes.CDLL(out)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


FORMS = [(0, 0.1, 36.0, 1.0, 1.0, 0.0), (0, 1.0 / 70, 36.0, 1.0, 1.0, 0.0), (1, 0.01, 36.0, 1.0, 1.0, 0.0), (1, 5.0, 36.0, 1.0, 1.0, 0.0),
         (2, 0.1, 36.0, 1.0, 1.0, 0.2), (2, 0.1, 36.0, 1.0, -1.0, 0.0)]


@pytest.mark.parametrize("gd,vdeg", [(3, 1), (3, 2), (2, 1), (2, 2)])
@pytest.mark.parametrize("form", FORMS)
def test_blocks_equal_rows(host, gd, vdeg, form):
    rng = np.random.default_rng(100 * gd + 10 * vdeg + form[0])
    nvn = gd + 1 if vdeg == 1 else (10 if gd == 3 else 6)
    nd = gd * nvn + gd + 1
    for trial in range(6):
        x = np.zeros((gd + 1, 3))
        x[1:, :gd] = np.eye(gd) * 0.3
        x[:, :gd] += rng.uniform(-0.08, 0.08, size=(gd + 1, gd)) + rng.uniform(-1, 1, size=(1, gd))
        if trial % 2:
            x[[0, 1]] = x[[1, 0]]                      # negative det J
        w = rng.normal(size=nd) * (1.0 if trial < 4 else 1e-3)
        Ab, Ar = np.empty((nd, nd)), np.empty((nd, nd))
        bb, br = np.empty(nd), np.empty(nd)
        rc = host.block_vs_rows(gd, vdeg, form[0], *[ctypes.c_double(v) for v in form[1:]], _p(x), _p(w), _p(Ab), _p(bb), _p(Ar), _p(br))
        assert rc == 0
        scale = np.abs(Ar).max()
        assert np.abs(Ab - Ar).max() <= 2e-13 * scale, (gd, vdeg, form, np.abs(Ab - Ar).max() / scale)
        assert np.abs(bb - br).max() <= 2e-13 * max(np.abs(br).max(), 1e-30)
        if form[0] == 0:   # the lean G-metric blocks (row-side records + mixed part), the row-owner kernel's flavour-0 path
            Al, bl = np.empty((nd, nd)), np.empty(nd)
            assert host.lean_blocks(gd, vdeg, ctypes.c_double(form[1]), ctypes.c_double(form[2]), _p(x), _p(w), _p(Al), _p(bl)) == 0
            assert np.abs(Al - Ar).max() <= 2e-13 * scale, (gd, vdeg, form, np.abs(Al - Ar).max() / scale)
            assert np.abs(bl - br).max() <= 2e-13 * max(np.abs(br).max(), 1e-30)
