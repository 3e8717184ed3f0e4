"""Reference-side pins on the GPU path (-m gpu): the two numbers the reference's own authors check against (tests/_pins.py).
The same DFG run on the CPU oracle is tests/test_oracle.py::test_oracle_reproduces_the_dfg_2d_1_reference_values."""
import numpy as np
import pytest
import scipy.sparse as sps
import scipy.sparse.linalg as spla

import _pins as P
from stabilized_navier_stokes_flow_fenicsx_b200 import mesh as M
from stabilized_navier_stokes_flow_fenicsx_b200.assembler import NSAssembler

pytestmark = pytest.mark.gpu


def test_dfg_2d_1_drag_and_lift_match_the_reference_constants():
    """DFG_2D_Validation.py:202-203 compares against Cd = 5.57953523384 and Cl = 0.010618948146 (Schaefer-Turek 2D-1, Re = 20).
    UGN-stabilised P1-P1 triangles assembled by libnsgpu, Newton through the NonlinearProblem adapter, 5 462 cells."""
    m, sp, bcs, obstacle = P.dfg_problem()
    be = P.GpuBackend(m, sp, bcs)
    r = P.dfg_solve(be, sp, bcs, obstacle)
    be.close()
    assert r["newton"][-1][1] < 1e-8 and len(r["newton"]) <= 8, r["newton"]
    assert abs(r["cd"] / P.DFG_CD - 1.0) < 0.01, r["cd"]
    assert abs(r["cl"] / P.DFG_CL - 1.0) < 0.05, r["cl"]
    assert abs(r["dp"] / P.DFG_DP - 1.0) < 0.04, r["dp"]


def test_duct_stokes_outlet_is_the_fully_developed_square_duct_profile():
    """StokesFlow/DuctStokesFlow.py:188-203 ("has a known output", README.md:43-56): P2-P1 Taylor-Hood Stokes flow in the square duct
    of length 4 with a uniform inlet; at the outlet the centreline velocity is 2.0962 x the mean (series solution).  The
    operator is assembled by libnsgpu (flavour STOKES with the script's pressure sign), solved once with a sparse LU."""
    m = M.duct_mesh(6, 24); sp = M.mixed_space(m, 2)
    bcs = P.uniform_inlet_duct_bcs(sp)
    asm = NSAssembler(m.x, m.cells, sp.dofmap, vdeg=2)
    asm.set_form(flavour=2, nu=1.0, alpha=1.0, sp=-1.0, beta=0.0)
    asm.set_bcs(bcs)
    indptr, indices = asm.create_matrix()
    marker = np.zeros(sp.n_dofs, dtype=bool); value = np.zeros(sp.n_dofs)
    for d, v in bcs:
        marker[d] = True; value[d] = v
    w = np.where(marker, value, 0.0)
    vals, F = asm.jacobian_residual(w)
    asm.close()
    A = sps.csr_matrix((vals, indices, indptr), shape=(sp.n_dofs, sp.n_dofs)).tocsc()
    w = w - spla.spsolve(A, F)
    influx, area = P.plane_flux(m, sp, w, 0.0)
    outflux, area_o = P.plane_flux(m, sp, w, 4.0)
    assert abs(area - 1.0) < 1e-12 and abs(area_o - 1.0) < 1e-12
    # Taylor-Hood conserves mass against every pressure test function; the outlet pressure dofs are Dirichlet rows (p = 0), so
    # q = 1 is not one of them and the net flux vanishes only up to the divergence in the last cell layer (oracle: -2.0e-4)
    assert abs(outflux - influx) < 5e-4 * abs(influx)
    X, c = sp.dof_x, sp.dof_comp
    ctr = np.flatnonzero((c == 0) & (np.abs(X[:, 0] - 4.0) < 1e-12) & (np.abs(X[:, 1]) < 1e-12) & (np.abs(X[:, 2]) < 1e-12))
    ratio = w[ctr[0]] / (outflux / area_o)
    assert abs(ratio / P.DUCT_RATIO - 1.0) < 5e-3, ratio
    # the profile is fully developed: the same ratio one diameter upstream, no cross-flow at the outlet
    ctr_up = np.flatnonzero((c == 0) & (np.abs(X[:, 0] - 3.0) < 1e-12) & (np.abs(X[:, 1]) < 1e-12) & (np.abs(X[:, 2]) < 1e-12))
    assert abs(w[ctr_up[0]] / w[ctr[0]] - 1.0) < 2e-3
    out_v = (c > 0) & (c < 3) & (np.abs(X[:, 0] - 4.0) < 1e-12)
    assert np.abs(w[out_v]).max() < 2e-3


@pytest.mark.parametrize("n_cross,n_long,tol,pc", [(10, 40, 0.04, 4), (16, 64, 0.02, 5)])
def test_duct_navier_stokes_on_the_device_converges_to_the_duct_profile(n_cross, n_long, tol, pc):
    """The flagship path end to end on the device (NavierStokesChannelFlow.py:268-293): G-metric P1-P1 assembly, KSPTFQMR with the
    4x4 vertex-block Jacobi (pc = 4) or the multicolour block ILU(0) (pc = 5), Newton with the bt line search, nothing but scalars over PCIe.  Re = 10 in the duct of length 4:
    the outlet profile is the developed square-duct one (ratio 2.0962), approached under refinement."""
    m = M.duct_mesh(n_cross, n_long); sp = M.mixed_space(m, 1)
    bcs = M.duct_bcs(sp)
    asm = NSAssembler(m.x, m.cells, sp.dofmap, vdeg=1)
    asm.set_form(flavour=0, nu=0.1); asm.set_bcs(bcs)
    asm.create_matrix(fetch=False)
    marker = np.zeros(sp.n_dofs, dtype=bool); value = np.zeros(sp.n_dofs)
    for d, v in bcs:
        marker[d] = True; value[d] = v
    w = np.where(marker, value, 0.0)
    w_dev = asm.dev_alloc(8 * asm.n_cols)
    asm.h2d(w_dev, w)
    hist = asm.newton_dev(w_dev, rtol=1e-8, atol=1e-8, max_it=30, ksp_rtol=1e-8, ksp_max_it=5000, pc=pc, linesearch="bt")
    asm.d2h(w, w_dev)
    assert asm.last_kernel_name() == "p1tet_ws"
    asm.close()
    assert hist[-1]["fnorm"] <= max(1e-8, 1e-8 * hist[0]["fnorm"]), hist
    influx, _ = P.plane_flux(m, sp, w, 0.0)
    outflux, area = P.plane_flux(m, sp, w, 4.0)
    # PSPG-stabilised P1-P1 is not exactly mass conserving: the flux defect is O(h^2) (measured -1.3 % at 10x10x40, -0.5 % at 16x16x64)
    assert abs(outflux - influx) < 0.5 * tol * abs(influx)
    if pc == 5:
        assert sum(h.get("ksp_its", 0) for h in hist) < 1000, hist      # block Jacobi needs ~6 700 Krylov iterations on this mesh
    X, c = sp.dof_x, sp.dof_comp
    ctr = np.flatnonzero((c == 0) & (np.abs(X[:, 0] - 4.0) < 1e-12) & (np.abs(X[:, 1]) < 1e-12) & (np.abs(X[:, 2]) < 1e-12))
    ratio = w[ctr[0]] / (outflux / area)
    assert abs(ratio / P.DUCT_RATIO - 1.0) < tol, (ratio, hist)
