"""oracle.py -- ctypes loader + NumPy helpers for the CPU oracle (TEST INFRASTRUCTURE).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  The product package never does.  PARITY UNPINNED against dolfinx itself (not
installable here) -- see the header of ns_oracle.c.

Global-assembly semantics restated from the reference call sites
NavierStokes/NavierStokesChannelFlow.py:51-75 (F / J callbacks), :271-272 (create_matrix /
create_petsc_vector) and SURVEY.md Appendix A.5.
"""
import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

GMETRIC, UGN, STOKES = 0, 1, 2


@dataclass
class Form:
    """Which weak form, on which element pair.  flavour: 0 G-metric NS, 1 UGN NS, 2 Stokes."""
    flavour: int = GMETRIC
    gdim: int = 3
    vdeg: int = 1
    nu: float = 0.1
    Ci: float = 36.0
    alpha: float = 1.0
    sp: float = 1.0
    beta: float = 0.0

    @property
    def ndofs_cell(self):
        nvn = self.gdim + 1 if self.vdeg == 1 else (10 if self.gdim == 3 else 6)
        return self.gdim * nvn + self.gdim + 1

    def args(self):
        return (ctypes.c_int(self.flavour), ctypes.c_int(self.gdim), ctypes.c_int(self.vdeg), ctypes.c_double(self.nu),
                ctypes.c_double(self.Ci), ctypes.c_double(self.alpha), ctypes.c_double(self.sp), ctypes.c_double(self.beta))


def build():
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libns_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.oracle_num_threads.restype = ctypes.c_int
    return _LIB


def _p(a, t=None):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


def element(form, x, w, want_A=True, want_b=True):
    """Element Jacobian Ae (nd x nd, row-major test x trial) and residual be (nd) for one cell."""
    nd = form.ndofs_cell
    x = np.ascontiguousarray(x, dtype=np.float64).reshape(form.gdim + 1, 3)
    w = np.ascontiguousarray(w, dtype=np.float64)
    Ae = np.zeros((nd, nd)) if want_A else None
    be = np.zeros(nd) if want_b else None
    lib().oracle_element_py(*form.args(), _p(x), _p(w), _p(Ae), _p(be))
    return Ae, be


def build_pattern(dofmap, n_rows, n_cells=None):
    """CSR sparsity = sorted-unique union over cells of dofs x dofs (create_matrix(problem.a),
    NavierStokesChannelFlow.py:272).  int64 indptr, int32 sorted column indices."""
    dm = np.asarray(dofmap)[: n_cells if n_cells is not None else len(dofmap)].astype(np.int64)
    nd = dm.shape[1]
    ncols = int(dm.max()) + 1 if dm.size else 1
    rows = np.repeat(dm, nd, axis=1).ravel()
    cols = np.tile(dm, (1, nd)).ravel()
    keys = np.unique(rows * ncols + cols)
    r, c = keys // ncols, keys % ncols
    indptr = np.zeros(n_rows + 1, dtype=np.int64)
    np.add.at(indptr, r + 1, 1)
    return np.cumsum(indptr), c.astype(np.int32)


def build_pattern_c(dofmap, n_rows, n_cells=None):
    """Same pattern as build_pattern, computed row by row in C (OpenMP): for the meshes of millions of cells where the
    NumPy set union needs tens of GB."""
    dm = np.ascontiguousarray(np.asarray(dofmap)[: n_cells if n_cells is not None else len(dofmap)], dtype=np.int32)
    indptr = np.zeros(n_rows + 1, dtype=np.int64)
    f = lib().oracle_build_pattern
    f.restype = ctypes.c_int64
    nnz = f(ctypes.c_int(dm.shape[1]), ctypes.c_int64(dm.shape[0]), _p(dm), ctypes.c_int64(n_rows), _p(indptr), None)
    if nnz < 0:
        raise RuntimeError("oracle_build_pattern failed")
    indices = np.empty(nnz, dtype=np.int32)
    if f(ctypes.c_int(dm.shape[1]), ctypes.c_int64(dm.shape[0]), _p(dm), ctypes.c_int64(n_rows), _p(indptr), _p(indices)) != nnz:
        raise RuntimeError("oracle_build_pattern failed")
    return indptr, indices


def assemble_residual(form, x, cells, dofmap, w, bc_marker=None, bc_value=None, n_cells_owned=None, lifting=True):
    """assemble_vector(F, L) + apply_lifting(F, [a], [bc], [x], -1.0)  (:64-65).  No set_bc, no halo."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    cells = np.ascontiguousarray(cells, dtype=np.int32)
    dofmap = np.ascontiguousarray(dofmap, dtype=np.int32)
    w = np.ascontiguousarray(w, dtype=np.float64)
    nc = len(cells) if n_cells_owned is None else n_cells_owned
    b = np.zeros_like(w)
    m = None if bc_marker is None else np.ascontiguousarray(bc_marker, dtype=np.uint8)
    g = None if bc_value is None else np.ascontiguousarray(bc_value, dtype=np.float64)
    lib().oracle_assemble_residual(*form.args(), _p(x), _p(cells), _p(dofmap), ctypes.c_int64(nc), _p(w), _p(m), _p(g),
                                   ctypes.c_int(1 if (lifting and m is not None) else 0), _p(b))
    return b


def set_bc(b, bc_dofs_list, bc_vals_list, x0, alpha=-1.0, n_owned=None):
    """set_bc(F, bcs, x, -1.0) (:67):  b[dof] = alpha * (g - x0[dof]) for owned BC dofs, list order."""
    for dofs, vals in zip(bc_dofs_list, bc_vals_list):
        dofs = np.asarray(dofs)
        vals = np.asarray(vals, dtype=np.float64)
        if n_owned is not None:
            keep = dofs < n_owned
            dofs, vals = dofs[keep], vals[keep]
        b[dofs] = alpha * (vals - x0[dofs])
    return b


def bc_arrays(n_dofs, bc_dofs_list, bc_vals_list):
    """marker / value / multiplicity arrays from a list of DirichletBC-like (dofs, values) pairs.
    Value: last BC in list order wins (dolfinx applies them in order)."""
    marker = np.zeros(n_dofs, dtype=np.uint8)
    value = np.zeros(n_dofs)
    mult = np.zeros(n_dofs, dtype=np.int32)
    for dofs, vals in zip(bc_dofs_list, bc_vals_list):
        dofs = np.asarray(dofs)
        marker[dofs] = 1
        value[dofs] = np.asarray(vals, dtype=np.float64)
        np.add.at(mult, dofs, 1)
    return marker, value, mult


def assemble_jacobian(form, x, cells, dofmap, w, indptr, indices, bc_marker=None, bc_mult=None, n_owned=None,
                      n_cells_owned=None):
    """J.zeroEntries(); assemble_matrix(J, a, bcs=bc) incl. per-BC-object unit diagonals (:73-74)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    cells = np.ascontiguousarray(cells, dtype=np.int32)
    dofmap = np.ascontiguousarray(dofmap, dtype=np.int32)
    w = np.ascontiguousarray(w, dtype=np.float64)
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    nc = len(cells) if n_cells_owned is None else n_cells_owned
    vals = np.zeros(len(indices))
    m = None if bc_marker is None else np.ascontiguousarray(bc_marker, dtype=np.uint8)
    mu = None if bc_mult is None else np.ascontiguousarray(bc_mult, dtype=np.int32)
    n_owned = len(indptr) - 1 if n_owned is None else n_owned
    err = lib().oracle_assemble_jacobian(*form.args(), _p(x), _p(cells), _p(dofmap), ctypes.c_int64(nc), _p(w), _p(m), _p(mu),
                                         ctypes.c_int64(n_owned), _p(indptr), _p(indices), _p(vals))
    if err:
        raise RuntimeError("oracle: entry outside the sparsity pattern")
    return vals


def spmv(indptr, indices, vals, xv):
    indptr = np.ascontiguousarray(indptr, dtype=np.int64)
    indices = np.ascontiguousarray(indices, dtype=np.int32)
    vals = np.ascontiguousarray(vals, dtype=np.float64)
    xv = np.ascontiguousarray(xv, dtype=np.float64)
    n = len(indptr) - 1
    y = np.zeros(n)
    lib().oracle_spmv(ctypes.c_int64(n), _p(indptr), _p(indices), _p(vals), _p(xv), _p(y))
    return y


def set_num_threads(n):
    """OpenMP team size of the oracle's loops (torchrun exports OMP_NUM_THREADS=1; the CPU baseline wants all cores)."""
    lib().oracle_set_num_threads(ctypes.c_int(int(n)))


def num_threads():
    return int(lib().oracle_num_threads())
