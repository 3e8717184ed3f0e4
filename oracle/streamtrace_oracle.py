"""TEST INFRASTRUCTURE ONLY -- CPU restatement of the reference's streamline tracing (NavierStokes/streamtrace.py).

The integrator IS the reference's: scipy.integrate.solve_ivp(method='RK45', events=..., max_step=0.125), called exactly
as streamtrace_pool (:198-218) and reverse_streamtrace_pool (:357-384) call it, with the same event functions
(:177-190).  What is restated is the right-hand side velfunc (:144-158): dolfinx's bb_tree /
compute_colliding_cells / uh.eval are replaced by a brute-force search over all cells (lowest-numbered cell whose
barycentric coordinates are all >= -tol) and the P1 interpolation  sum_a lambda_a(x) u_a; outside the mesh the velocity
is zero.  scipy is the pinned third-party algorithm here (environment.yml: scipy), so the pin is the real one.
"""
import numpy as np
from scipy.integrate import solve_ivp


class TraceOracle:
    def __init__(self, x, cells, u, tol=1e-12):
        self.x = np.asarray(x, dtype=np.float64)
        self.cells = np.asarray(cells, dtype=np.int64)
        self.u = np.asarray(u, dtype=np.float64).reshape(-1, 3)
        self.tol = tol
        x0 = self.x[self.cells[:, 0]]
        J = np.stack([self.x[self.cells[:, a + 1]] - x0 for a in range(3)], axis=2)   # J[c, i, a]
        self.x0 = x0
        self.K = np.linalg.inv(J)                                                   # K[c, a, i]: gradient of lambda_{a+1}

    def locate(self, p):
        lam = np.einsum("cai,ci->ca", self.K, p[None, :] - self.x0)
        l0 = 1.0 - lam.sum(axis=1)
        ok = (lam.min(axis=1) >= -self.tol) & (l0 >= -self.tol)
        idx = np.flatnonzero(ok)
        if idx.size == 0:
            return -1, None
        c = int(idx[0])
        return c, np.concatenate(([l0[c]], lam[c]))

    def velfunc(self, t, p):                       # streamtrace.py:144-158
        c, lam = self.locate(np.asarray(p, dtype=np.float64))
        if c < 0:
            return np.array([0.0, 0.0, 0.0])
        return lam @ self.u[self.cells[c]]

    def velfunc_reverse(self, t, p):               # :160-174
        return -self.velfunc(t, p)

    def forward(self, seed, x_stop=3.7, speed_min=1e-6, t_end=20.0, max_step=0.125):    # :198-218
        def velocity_magnitude_event(t, y):
            return np.linalg.norm(self.velfunc(t, y)) - speed_min
        def position_event(t, y):
            return y[0] - x_stop
        velocity_magnitude_event.terminal = True
        velocity_magnitude_event.direction = -1
        position_event.terminal = True
        position_event.direction = 1
        return solve_ivp(self.velfunc, (0, t_end), seed, method="RK45", events=(velocity_magnitude_event, position_event), max_step=max_step)

    def reverse(self, seed, x_stop=0.13, speed_min=1e-6, t_end=20.0, max_step=0.125):   # :357-384
        def velocity_magnitude_event(t, y):
            return np.linalg.norm(self.velfunc(t, y)) - speed_min
        def reverse_position_event(t, y):
            return y[0] - x_stop
        velocity_magnitude_event.terminal = True
        velocity_magnitude_event.direction = -1
        reverse_position_event.terminal = True
        reverse_position_event.direction = -1
        return solve_ivp(self.velfunc_reverse, (0, t_end), seed, method="RK45", events=(reverse_position_event, velocity_magnitude_event), max_step=max_step)


def host_tables(x, cells, u, tol=1e-12):
    """The locator / velocity tables of csrc/streamtrace.cu built with NumPy (for the g++ harness of trace_core.cuh)."""
    x = np.asarray(x, dtype=np.float64)
    cells = np.asarray(cells, dtype=np.int64)
    u = np.asarray(u, dtype=np.float64).reshape(-1, 3)
    nc = cells.shape[0]
    x0 = x[cells[:, 0]]
    J = np.stack([x[cells[:, a + 1]] - x0 for a in range(3)], axis=2)
    K = np.linalg.inv(J)
    cmap = np.concatenate([x0, K.reshape(nc, 9)], axis=1)
    du = np.stack([u[cells[:, a + 1]] - u[cells[:, 0]] for a in range(3)], axis=1)      # du[c, a, i]
    A = np.einsum("cai,caj->cij", du, K)
    c0 = u[cells[:, 0]] - np.einsum("cij,cj->ci", A, x0)
    cvel = np.concatenate([c0, A.reshape(nc, 9)], axis=1)
    lo, hi = x.min(axis=0), x.max(axis=0)
    ext = hi - lo
    s = 1.5 * np.cbrt(np.prod(np.where(ext > 0, ext, 1.0)) / nc)
    nb = np.clip(np.ceil(ext / s).astype(np.int64), 1, 1024)
    inv_h = np.where(ext > 0, nb / np.where(ext > 0, ext, 1.0), 1.0)
    pts = x[cells]                                                                        # (nc, 4, 3)
    pad = 1e-9 / inv_h
    b0 = np.clip(np.floor((pts.min(axis=1) - pad - lo) * inv_h).astype(np.int64), 0, nb - 1)
    b1 = np.clip(np.floor((pts.max(axis=1) + pad - lo) * inv_h).astype(np.int64), 0, nb - 1)
    keys = []
    for c in range(nc):
        kk, jj, ii = np.meshgrid(np.arange(b0[c, 2], b1[c, 2] + 1), np.arange(b0[c, 1], b1[c, 1] + 1), np.arange(b0[c, 0], b1[c, 0] + 1), indexing="ij")
        bins = (kk * nb[1] + jj) * nb[0] + ii
        keys.append((bins.ravel() << 32) | c)
    keys = np.sort(np.concatenate(keys))
    n_bins = int(np.prod(nb))
    bin_ptr = np.searchsorted(keys, np.arange(n_bins + 1, dtype=np.int64) << 32).astype(np.int64)
    bin_cells = (keys & 0xFFFFFFFF).astype(np.int32)
    return dict(cmap=np.ascontiguousarray(cmap), cvel=np.ascontiguousarray(cvel), bin_ptr=bin_ptr, bin_cells=bin_cells,
                lo=lo.astype(np.float64), inv_h=inv_h.astype(np.float64), nb=nb.astype(np.int32), tol=tol)
