"""Device-resident TFQMR (KSPTFQMR, NavierStokes/NavierStokesChannelFlow.py:77, :282-285) and the Newton loop built on it.

The linear solve is checked against a sparse direct solve of the same (GPU-assembled, oracle-checked) Jacobian; the Newton
loop -- assembly kernels + TFQMR + axpy, everything on the device -- against Newton with the oracle's F / J and a direct
solve on the host: converged velocity / pressure within 1e-8 relative L2 (the north-star tolerance)."""
import numpy as np
import pytest
import scipy.sparse as sps
import scipy.sparse.linalg as spla

from stabilized_navier_stokes_flow_fenicsx_b200 import mesh as M
from stabilized_navier_stokes_flow_fenicsx_b200.assembler import NSAssembler

pytestmark = pytest.mark.gpu


def _duct(nc=4, nl=10):
    m = M.duct_mesh(nc, nl)
    sp = M.mixed_space(m, 1)
    return m, sp, M.duct_bcs(sp), dict(flavour=0, nu=0.1)


@pytest.mark.parametrize("pc", [0, 1, 4])
def test_tfqmr_matches_direct_solve(pc):
    m, sp, bcs, fk = _duct()
    asm = NSAssembler(m.x, m.cells, sp.dofmap, vdeg=1)
    asm.set_form(**fk); asm.set_bcs(bcs)
    indptr, indices = asm.create_matrix()
    w = M.duct_state(sp)
    vals, F = asm.jacobian_residual(w)
    A = sps.csr_matrix((vals, indices, indptr), shape=(sp.n_dofs, sp.n_dofs))
    x_ref = spla.spsolve(A.tocsc(), F)
    x, info = asm.tfqmr(F, rtol=1e-12, max_it=4000, pc=pc)
    assert info["rnorm"] <= 1e-9 * info["r0norm"], info          # true residual, recomputed at exit
    assert np.linalg.norm(x - x_ref) <= 1e-8 * np.linalg.norm(x_ref), info
    # warm start from the solution: nothing left to do
    x2, info2 = asm.tfqmr(F, x0=x, rtol=1e-8, max_it=10, pc=pc)
    assert info2["its"] <= 1 and np.linalg.norm(x2 - x_ref) <= 1e-8 * np.linalg.norm(x_ref)
    asm.close()


def test_block_jacobi_needs_fewer_iterations_than_none():
    m, sp, bcs, fk = _duct(6, 16)
    asm = NSAssembler(m.x, m.cells, sp.dofmap, vdeg=1)
    asm.set_form(**fk); asm.set_bcs(bcs)
    asm.create_matrix(fetch=False)
    _, F = asm.jacobian_residual(M.duct_state(sp), fetch_vals=False)
    its = {pc: asm.tfqmr(F, rtol=1e-8, max_it=5000, pc=pc)[1]["its"] for pc in (0, 4)}
    assert its[4] < its[0], its
    asm.close()


def test_tfqmr_on_p2p1_falls_back_to_point_jacobi():
    m = M.duct_mesh(3, 6)
    sp = M.mixed_space(m, 2)
    asm = NSAssembler(m.x, m.cells, sp.dofmap, vdeg=2)
    asm.set_form(flavour=0, nu=0.02)                                # P2-P1 G-metric form (config 4): PSPG gives a pressure diagonal
    asm.set_bcs(M.duct_bcs(sp))
    indptr, indices = asm.create_matrix()
    vals, F = asm.jacobian_residual(M.duct_state(sp))
    A = sps.csr_matrix((vals, indices, indptr), shape=(sp.n_dofs, sp.n_dofs))
    b = A @ np.random.default_rng(3).standard_normal(sp.n_dofs)
    x, info = asm.tfqmr(b, rtol=1e-10, max_it=20000, pc=4)          # 4 is not applicable here: point Jacobi
    assert info["rnorm"] <= 1e-6 * info["r0norm"], info
    asm.close()


def test_device_newton_converges_to_the_oracle_fixed_point(oracle):
    m, sp, bcs, fk = _duct()
    form = oracle.Form(gdim=3, vdeg=1, **fk)
    marker, value, mult = oracle.bc_arrays(sp.n_dofs, [b[0] for b in bcs], [b[1] for b in bcs])
    indptr, indices = oracle.build_pattern(sp.dofmap, sp.n_dofs)
    w0 = np.zeros(sp.n_dofs)
    w0[marker == 1] = value[marker == 1]     # as the reference: the initial guess meets the Dirichlet values
    # host Newton on the oracle
    w = w0.copy()
    for _ in range(25):
        F = oracle.assemble_residual(form, m.x, m.cells, sp.dofmap, w, marker, value)
        F = oracle.set_bc(F, [b[0] for b in bcs], [b[1] for b in bcs], w)
        if np.linalg.norm(F) < 1e-12:
            break
        J = oracle.assemble_jacobian(form, m.x, m.cells, sp.dofmap, w, indptr, indices, marker, mult)
        w = w - spla.spsolve(sps.csr_matrix((J, indices, indptr), shape=(sp.n_dofs,) * 2).tocsc(), F)
    # device Newton: assembly kernels + TFQMR, state never leaves the GPU
    asm = NSAssembler(m.x, m.cells, sp.dofmap, vdeg=1)
    asm.set_form(**fk); asm.set_bcs(bcs)
    asm.create_matrix(fetch=False)
    w_dev = asm.dev_alloc(8 * asm.n_cols)
    asm.h2d(w_dev, w0)
    hist = asm.newton_dev(w_dev, rtol=0.0, atol=1e-11, max_it=30, ksp_rtol=1e-10, ksp_max_it=4000, pc=4)
    wg = np.zeros(asm.n_cols)
    asm.d2h(wg, w_dev)
    assert hist[-1]["fnorm"] <= 1e-11, hist
    vel = sp.dof_comp < 3
    for part in (vel, ~vel):
        assert np.linalg.norm(wg[: sp.n_dofs][part] - w[part]) <= 1e-8 * np.linalg.norm(w[part]), hist
    asm.dev_free(w_dev)
    asm.close()
