"""Direct parity against the reference's own stack (dolfinx 0.9 / PETSc): pattern bit-exact, entries and residual to 1e-12.

Skipped wherever dolfinx cannot be imported -- which includes the image this repository was built in, so the
oracle stays "parity unpinned" until this test has run somewhere; tools/compare_with_dolfinx.py is the same check as a script."""
import os
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


@pytest.mark.parametrize("vdeg", [1, 2])
def test_same_arrays_same_matrix_as_dolfinx(vdeg):
    pytest.importorskip("dolfinx")
    pytest.importorskip("petsc4py")
    import compare_with_dolfinx as C
    r = C.run(n=(4, 4, 8) if vdeg == 1 else (3, 3, 5), Re=10.0, vdeg=vdeg, verbose=False)
    assert r["pattern_equal"], r
    assert r["J_rel_err"] <= 1e-12 and r["F_rel_err"] <= 1e-12, r
