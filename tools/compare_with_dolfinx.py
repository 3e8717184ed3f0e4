#!/usr/bin/env python
"""tools/compare_with_dolfinx.py -- the pin this repository could not create in its build container.

Runs the reference's own assembly sequence (NavierStokes/NavierStokesChannelFlow.py:40-75: fem.form, create_matrix,
assemble_matrix with bcs, assemble_vector, apply_lifting(..., -1.0), set_bc(..., -1.0)) with dolfinx / PETSc on a small
tetrahedral box, hands the SAME arrays (mesh.geometry.x, mesh.geometry.dofmap, W.dofmap.list, one (dofs, values) pair per
dirichletbc) to libnsgpu, and compares:

  * CSR sparsity: indptr / indices bit-exact with A.getValuesCSR();
  * Jacobian entries and residual: max |difference| <= 1e-12 * max |reference|  (north-star tolerance).

It needs dolfinx 0.9 + petsc4py + a CUDA device; none of FEniCSx is installable in the image this repository was built in
(SURVEY.md 8c), so THIS SCRIPT HAS NOT BEEN EXECUTED THERE -- it is written against the dolfinx 0.9 Python API as the
reference uses it, and tests/test_dolfinx_parity.py runs it wherever `import dolfinx` succeeds.  Serial (one rank).

  python tools/compare_with_dolfinx.py [--n 6 6 12] [--re 10] [--p2]
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def weak_form(W, msh, Re):
    """The G-metric SUPG/PSPG/LSIC residual of define_navier_stokes_form (NavierStokesChannelFlow.py:220-251), restated."""
    import ufl
    from dolfinx.fem import Function
    dx = ufl.dx(metadata={"quadrature_degree": 2})
    nu = 1.0 / Re
    w = Function(W)
    u, p = ufl.split(w)
    v, q = ufl.TestFunctions(W)
    K = ufl.inv(ufl.Jacobian(msh)) * ufl.inv(ufl.grad(ufl.SpatialCoordinate(msh)))      # d xi / d x
    G = K.T * K
    tau = 1.0 / ufl.sqrt(ufl.inner(u, G * u) + 36.0 * nu**2 * ufl.inner(G, G))
    sigma = 2 * nu * ufl.sym(ufl.grad(u)) - p * ufl.Identity(len(u))
    r_m = ufl.dot(u, ufl.grad(u)) - ufl.div(sigma)
    F = ufl.inner(ufl.dot(u, ufl.nabla_grad(u)), v) * dx + nu * ufl.inner(ufl.grad(u), ufl.grad(v)) * dx
    F += -ufl.inner(p, ufl.div(v)) * dx + ufl.inner(q, ufl.div(u)) * dx
    F += ufl.inner(tau * r_m, ufl.dot(u, ufl.grad(v)) + ufl.grad(q)) * dx
    F += (1.0 / (ufl.tr(G) * tau)) * ufl.div(v) * ufl.div(u) * dx
    return F, w, ufl.derivative(F, w, ufl.TrialFunction(W))


def run(n=(6, 6, 12), Re=10.0, vdeg=1, verbose=True):
    from mpi4py import MPI
    import basix.ufl
    from dolfinx import fem, mesh
    from dolfinx.fem.petsc import apply_lifting, assemble_matrix, assemble_vector, create_matrix, create_vector, set_bc
    from petsc4py import PETSc
    from stabilized_navier_stokes_flow_fenicsx_b200.assembler import NSAssembler

    comm = MPI.COMM_SELF
    msh = mesh.create_box(comm, [np.array([0.0, -0.5, -0.5]), np.array([4.0, 0.5, 0.5])], [n[2], n[0], n[1]], mesh.CellType.tetrahedron)
    Ve = basix.ufl.element("Lagrange", msh.basix_cell(), vdeg, shape=(3,))
    Qe = basix.ufl.element("Lagrange", msh.basix_cell(), 1)
    W = fem.functionspace(msh, basix.ufl.mixed_element([Ve, Qe]))
    F, w, dF = weak_form(W, msh, Re)

    # state: smooth field + noise, as bench.py's duct state
    rng = np.random.default_rng(1234)
    w.sub(0).interpolate(lambda x: np.stack([1.5 * (1 - 4 * x[1]**2) * (1 - 4 * x[2]**2) * (1 + 0.1 * np.sin(2 * np.pi * x[0])),
                                             0.05 * np.sin(2 * np.pi * x[1]), 0.05 * np.sin(2 * np.pi * x[2])]))
    w.sub(1).interpolate(lambda x: 4.0 - x[0])
    w.x.array[:] += 1e-3 * rng.standard_normal(w.x.array.size)

    # bcs = [wall, inlet, outlet] in the style of :134-146 (collapsed-space functions on sub-spaces; a dof may sit in two objects)
    W0, W1 = W.sub(0), W.sub(1)
    V0, _ = W0.collapse()
    Q0, _ = W1.collapse()
    noslip = fem.Function(V0)
    inlet = fem.Function(V0)
    inlet.interpolate(lambda x: np.stack([1.5 * (1 - 4 * x[1]**2) * (1 - 4 * x[2]**2), 0 * x[0], 0 * x[0]]))
    on_wall = lambda x: np.isclose(np.abs(x[1]), 0.5) | np.isclose(np.abs(x[2]), 0.5)
    bcs = [fem.dirichletbc(noslip, fem.locate_dofs_geometrical((W0, V0), on_wall), W0),
           fem.dirichletbc(inlet, fem.locate_dofs_geometrical((W0, V0), lambda x: np.isclose(x[0], 0.0)), W0),
           fem.dirichletbc(fem.Function(Q0), fem.locate_dofs_geometrical((W1, Q0), lambda x: np.isclose(x[0], 4.0)), W1)]

    # ---- the reference's sequence (:51-75)
    a_form, L_form = fem.form(dF), fem.form(F)
    A = create_matrix(a_form)
    A.zeroEntries()
    assemble_matrix(A, a_form, bcs=bcs)
    A.assemble()
    b = create_vector(L_form)
    x = w.x.petsc_vec
    with b.localForm() as bl:
        bl.set(0.0)
    assemble_vector(b, L_form)
    apply_lifting(b, [a_form], [bcs], [x], -1.0)
    b.ghostUpdate(addv=PETSc.InsertMode.ADD, mode=PETSc.ScatterMode.REVERSE)
    set_bc(b, bcs, x, -1.0)
    ai, aj, av = A.getValuesCSR()

    # ---- the same arrays through libnsgpu
    ndofs = W.dofmap.index_map.size_local * W.dofmap.index_map_bs
    assert W.dofmap.index_map_bs == 1 and W.dofmap.bs == 1, "mixed space expected to have block size 1"
    asm = NSAssembler(msh.geometry.x, msh.geometry.dofmap, W.dofmap.list, vdeg=vdeg)
    asm.set_form(flavour=0, nu=1.0 / Re, Ci=36.0)
    gbc = []
    for bc in bcs:
        dofs = np.asarray(bc.dof_indices()[0], dtype=np.int32)
        vals = np.zeros(ndofs)
        fem.set_bc(vals, [bc])                      # vals[dofs] = g
        gbc.append((dofs, vals[dofs]))
    asm.set_bcs(gbc)
    indptr, indices = asm.create_matrix()
    vals, Fg = asm.jacobian_residual(np.array(w.x.array))
    kernel = asm.last_kernel_name()
    asm.close()

    out = {"cells": int(msh.topology.index_map(3).size_local), "dofs": int(ndofs), "nnz": int(len(aj)), "kernel": kernel,
           "pattern_equal": bool(np.array_equal(indptr, ai) and np.array_equal(indices, aj))}
    if out["pattern_equal"]:
        out["J_rel_err"] = float(np.abs(vals - av).max() / np.abs(av).max())
    out["F_rel_err"] = float(np.abs(Fg[:ndofs] - b.array[:ndofs]).max() / np.abs(b.array).max())
    out["ok"] = bool(out["pattern_equal"] and out.get("J_rel_err", 1.0) <= 1e-12 and out["F_rel_err"] <= 1e-12)
    if verbose:
        print(out, flush=True)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, nargs=3, default=[6, 6, 12], help="boxes across (y, z) and along (x)")
    ap.add_argument("--re", type=float, default=10.0)
    ap.add_argument("--p2", action="store_true", help="P2-P1 Taylor-Hood instead of stabilized P1-P1")
    args = ap.parse_args()
    r = run(tuple(args.n), args.re, 2 if args.p2 else 1)
    sys.exit(0 if r["ok"] else 1)
