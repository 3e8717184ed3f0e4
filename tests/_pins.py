"""Shared drivers of the reference-side pins (tests/test_reference_pins.py, tests/test_oracle.py).

The reference holds exactly two numbers its own authors check results against:
  * DFG 2D-1 drag / lift, NavierStokes/Validation_Flow/DFG_2D_Validation.py:202-203: Cd = 5.57953523384, Cl = 0.010618948146;
  * the "known output" of StokesFlow/DuctStokesFlow.py (README.md:43-56): a fully developed square-duct profile at the outlet,
    whose centreline-to-mean velocity ratio is 2.0963 (series solution of the Poisson problem on the square).
Both are reproduced here through the same assembly calls the parity tests exercise, with either back end (``Backend``)."""
import numpy as np
import scipy.sparse as sps
import scipy.sparse.linalg as spla

from stabilized_navier_stokes_flow_fenicsx_b200 import mesh as M

DFG_CD, DFG_CL, DFG_DP = 5.57953523384, 0.010618948146, 0.11752016697
DUCT_RATIO = 2.0962


class OracleBackend:
    def __init__(self, oracle, m, sp, bcs):
        self.o, self.m, self.sp, self.bcs = oracle, m, sp, bcs
        self.marker, self.value, self.mult = oracle.bc_arrays(sp.n_dofs, [b[0] for b in bcs], [np.broadcast_to(b[1], (len(b[0]),)) for b in bcs])
        self.pattern = oracle.build_pattern_c(sp.dofmap, sp.n_dofs)

    def FJ(self, fk, w):
        form = self.o.Form(gdim=self.m.gdim, vdeg=self.sp.vdeg, **fk)
        F = self.o.assemble_residual(form, self.m.x, self.m.cells, self.sp.dofmap, w, self.marker, self.value)
        on = self.marker != 0
        F[on] = w[on] - self.value[on]
        J = self.o.assemble_jacobian(form, self.m.x, self.m.cells, self.sp.dofmap, w, self.pattern[0], self.pattern[1], self.marker, self.mult)
        return F, J

    def unconstrained_residual(self, fk, w):
        form = self.o.Form(gdim=self.m.gdim, vdeg=self.sp.vdeg, **fk)
        return self.o.assemble_residual(form, self.m.x, self.m.cells, self.sp.dofmap, w, None, None)

    def close(self):
        pass


class GpuBackend:
    """The C-ABI path: NSAssembler behind the NonlinearProblem-shaped adapter (LidDrivenNavierStokesFlow.py:150-151)."""
    def __init__(self, m, sp, bcs):
        from stabilized_navier_stokes_flow_fenicsx_b200.assembler import NSAssembler, NonlinearProblem
        self.m, self.sp, self.bcs = m, sp, bcs
        self.asm = NSAssembler(m.x, m.cells, sp.dofmap, vdeg=sp.vdeg)
        self.asm.set_bcs(bcs)
        self.problem = NonlinearProblem(self.asm)
        self.fk = None
        self.pattern = None

    def _form(self, fk):
        if fk != self.fk:
            self.asm.set_form(**fk)
            self.fk = dict(fk)
        if self.pattern is None:
            self.pattern = self.problem.create_matrix()

    def FJ(self, fk, w):
        self._form(fk)
        b, A = np.zeros(self.asm.n_dofs), np.zeros(self.asm.nnz)
        self.problem.form(w); self.problem.F(w, b); self.problem.J(w, A)
        return b, A

    def unconstrained_residual(self, fk, w):
        self._form(fk)
        self.asm.set_bcs([])
        R = self.asm.residual(w)
        self.asm.set_bcs(self.bcs)
        return R

    def close(self):
        self.asm.close()


def newton(be, fk, w, rtol=1e-9, max_it=30):
    """NewtonSolver with the incremental criterion and an LU solve (LidDrivenNavierStokesFlow.py:152-169)."""
    n = len(w)
    hist = []
    for _ in range(max_it):
        F, J = be.FJ(fk, w)
        A = sps.csr_matrix((J, be.pattern[1], be.pattern[0]), shape=(n, n)).tocsc()
        dx = spla.spsolve(A, F)
        w = w - dx
        hist.append((float(np.linalg.norm(F)), float(np.linalg.norm(dx))))
        if hist[-1][1] < rtol * max(1.0, float(np.linalg.norm(w))):
            break
    return w, hist


def dfg_problem(h_far=0.02, n_cyl=96):
    m = M.dfg_cylinder_mesh(h_far, n_cyl)
    sp = M.mixed_space(m, 1)
    bcs, obstacle = M.dfg_bcs(sp)
    return m, sp, bcs, obstacle


def dfg_solve(be, sp, bcs, obstacle, nu=1e-3):
    """DFG_2D_Validation.py: Stokes initial guess (:99-126), UGN-stabilised Navier-Stokes Newton solve (:134-190), drag and lift
    (:193-206) -- here as reaction forces: minus the unconstrained momentum residual summed over the cylinder's velocity dofs.
    The Stokes guess uses nu * grad u (the script's unit-viscosity guess makes the first full Newton step overshoot)."""
    marker = np.zeros(sp.n_dofs, dtype=bool); value = np.zeros(sp.n_dofs)
    for d, v in bcs:
        marker[d] = True; value[d] = v
    w = np.where(marker, value, 0.0)
    w, h_st = newton(be, dict(flavour=2, nu=nu, alpha=nu, sp=1.0, beta=0.2 / nu), w, max_it=3)
    fk = dict(flavour=1, nu=nu)
    w, h_ns = newton(be, fk, w)
    R = be.unconstrained_residual(fk, w)
    fx = -R[obstacle[sp.dof_comp[obstacle] == 0]].sum()
    fy = -R[obstacle[sp.dof_comp[obstacle] == 1]].sum()
    scale = 2.0 / (0.1 * 0.2 ** 2)                                # 2 / (D U_mean^2), DFG_2D_Validation.py:197-198
    from scipy.spatial import cKDTree
    pv = np.flatnonzero(sp.dof_comp == 2)
    tree = cKDTree(sp.dof_x[pv][:, :2])
    dp = w[pv[tree.query([0.15, 0.2])[1]]] - w[pv[tree.query([0.25, 0.2])[1]]]
    return dict(cd=scale * fx, cl=scale * fy, dp=dp, newton=h_ns, stokes=h_st, w=w)


def uniform_inlet_duct_bcs(space, length=4.0, tol=1e-12):
    """bcs = [bc_wall, bc_inlet, bc_outlet] of StokesFlow/DuctStokesFlow.py:149-183: no-slip walls, inlet velocity (1, 0, 0), p = 0 at the outlet."""
    X, c = space.dof_x, space.dof_comp
    vel = c < 3
    wall = np.flatnonzero(vel & ((np.abs(np.abs(X[:, 1]) - 0.5) < tol) | (np.abs(np.abs(X[:, 2]) - 0.5) < tol)))
    inlet = np.flatnonzero(vel & (np.abs(X[:, 0]) < tol))
    outlet = np.flatnonzero((c == 3) & (np.abs(X[:, 0] - length) < tol))
    return [(wall.astype(np.int32), np.zeros(len(wall))), (inlet.astype(np.int32), np.where(c[inlet] == 0, 1.0, 0.0)),
            (outlet.astype(np.int32), np.zeros(len(outlet)))]


def plane_flux(m, sp, w, xpos, tol=1e-12):
    """Exact integral of the finite-element u_x over the mesh faces lying in the plane x = xpos, and their total area."""
    cells = m.cells
    on = np.abs(m.x[:, 0] - xpos) < tol
    nv = m.n_vertices
    flux = area = 0.0
    edge_id = None
    if sp.vdeg == 2:
        key = np.minimum(sp.edges[:, 0], sp.edges[:, 1]).astype(np.int64) * nv + np.maximum(sp.edges[:, 0], sp.edges[:, 1])
        order = np.argsort(key)
        edge_id = (key[order], order)
    for f in ((0, 1, 2), (0, 1, 3), (0, 2, 3), (1, 2, 3)):
        tri = cells[:, f]
        sel = on[tri].all(axis=1)
        t = tri[sel].astype(np.int64)
        if not len(t):
            continue
        a, b, c = m.x[t[:, 0]], m.x[t[:, 1]], m.x[t[:, 2]]
        ar = 0.5 * np.linalg.norm(np.cross(b - a, c - a), axis=1)
        area += ar.sum()
        if sp.vdeg == 1:
            flux += (ar / 3.0 * (w[4 * t[:, 0]] + w[4 * t[:, 1]] + w[4 * t[:, 2]])).sum()          # P1: vertex rule is exact
        else:
            s = 0.0
            for i, j in ((0, 1), (1, 2), (0, 2)):                                                   # P2: edge-midpoint rule is exact
                k = np.minimum(t[:, i], t[:, j]) * nv + np.maximum(t[:, i], t[:, j])
                e = edge_id[1][np.searchsorted(edge_id[0], k)]
                s = s + w[4 * nv + 3 * e]
            flux += (ar / 3.0 * s).sum()
    return flux, area
