"""Host-side mirror of NavierStokes/streamtrace.py's tracing functions on top of the C ABI (nsgpu_trace_*).

``StreamTracer`` plays the role of the (bb_tree, mesh, uh) triple the reference passes around: it is built from the
arrays dolfinx exposes (``mesh.geometry.x``, ``mesh.geometry.dofmap`` of a tetrahedral mesh and the nodal values of the
P1 vector Function read by read_mesh_and_function, streamtrace.py:57-129).  ``run_streamtrace`` and
``run_reverse_streamtrace`` keep the names, argument meaning and return conventions of :220-250 and :386-446:
the forward trace returns the end points of the seeds that got past x = 0.5, the reverse trace one row per seed with
the (10, 10, 10) sentinel for seeds that did not come back to the inlet.  All seeds of a call are integrated at once on
the GPU (one thread per seed); there is no CPU fallback.
"""
import ctypes

import numpy as np

from . import _lib

REACHED_T_END, EVENT_POSITION, EVENT_SPEED, STEP_TOO_SMALL, MAX_STEPS = 0, 1, 2, -1, -2


def _ptr(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


class StreamTracer:
    def __init__(self, x, cells, u=None, device=0, tol=1e-12):
        self.lib = _lib.load()
        self.ctx = ctypes.c_void_p()
        rc = self.lib.nsgpu_create(ctypes.byref(self.ctx), device)
        if rc != 0:
            raise _lib.NsgpuError(f"nsgpu_create failed ({rc}): {self.lib.nsgpu_last_error(None).decode()}")
        x = np.ascontiguousarray(x, dtype=np.float64)
        cells = np.ascontiguousarray(cells, dtype=np.int32)
        if x.ndim != 2 or x.shape[1] != 3 or cells.ndim != 2 or cells.shape[1] != 4:
            raise ValueError("x must be (n_nodes, 3) and cells (n_cells, 4): a tetrahedral mesh like mesh.geometry.x / .dofmap")
        self.n_nodes = x.shape[0]
        self.tol = float(tol)
        _lib.check(self.ctx, self.lib.nsgpu_set_mesh(self.ctx, 3, x.shape[0], _ptr(x), cells.shape[0], cells.shape[0], _ptr(cells)), "set_mesh")
        self.set_velocity(u)

    def set_velocity(self, u):
        """(Re)load the nodal velocity (n_nodes x 3, geometry-node order); None only builds the locator."""
        if u is not None:
            u = np.ascontiguousarray(u, dtype=np.float64).reshape(-1, 3)
            if u.shape[0] != self.n_nodes:
                raise ValueError(f"velocity needs one row per geometry node ({self.n_nodes}), got {u.shape[0]}")
        _lib.check(self.ctx, self.lib.nsgpu_trace_setup(self.ctx, _ptr(u), self.tol), "trace_setup")

    def close(self):
        if getattr(self, "ctx", None):
            self.lib.nsgpu_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def velfunc(self, points, return_cells=False):
        """streamtrace.py:144-158 for an (n, 3) array of points: velocities (n, 3), zero outside the mesh."""
        p = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        vel = np.empty_like(p)
        cell = np.empty(p.shape[0], dtype=np.int32)
        _lib.check(self.ctx, self.lib.nsgpu_trace_velocity(self.ctx, p.shape[0], _ptr(p), _ptr(vel), _ptr(cell)), "trace_velocity")
        return (vel, cell) if return_cells else vel

    def trace(self, seeds, reverse=False, x_stop=None, speed_min=1e-6, t_span=(0, 20), max_step=0.125, rtol=1e-3, atol=1e-6,
              max_steps=1_000_000):
        """All seeds through solve_ivp(..., method='RK45', events=..., max_step=max_step): end points, status, final times, steps."""
        s = np.ascontiguousarray(seeds, dtype=np.float64).reshape(-1, 3)
        if t_span[0] != 0:
            raise ValueError("the field is autonomous: start the time span at 0")
        if x_stop is None:
            x_stop = 0.13 if reverse else 3.7          # streamtrace.py:188 / :183
        n = s.shape[0]
        end = np.empty((n, 3))
        status = np.empty(n, dtype=np.int32)
        t_final = np.empty(n)
        n_steps = np.empty(n, dtype=np.int32)
        _lib.check(self.ctx, self.lib.nsgpu_trace_run(self.ctx, n, _ptr(s), int(bool(reverse)), float(x_stop), float(speed_min), float(t_span[1]),
                                                      float(max_step), float(rtol), float(atol), int(max_steps), _ptr(end), _ptr(status),
                                                      _ptr(t_final), _ptr(n_steps)), "trace_run")
        return end, status, t_final, n_steps

    def last_kernel_ms(self):
        ms = (ctypes.c_double * 8)()
        _lib.check(self.ctx, self.lib.nsgpu_timers(self.ctx, ms, 8), "timers")
        return ms[7]


def run_streamtrace(inner_mesh, tracer):
    """streamtrace.py:220-250: forward-trace every row of inner_mesh; keep the end points with x > 0.5.
    Returns pointsx, pointsy, pointsz, each of shape (n_kept, 1) like the reference's np.array of one-element lists."""
    end, _, _, _ = tracer.trace(inner_mesh, reverse=False)
    keep = end[:, 0] > 0.5
    return end[keep, 0:1].copy(), end[keep, 1:2].copy(), end[keep, 2:3].copy()


def run_reverse_streamtrace(seeds, tracer):
    """streamtrace.py:386-446: reverse-trace every seed; seeds that do not come back below x = 0.5 get (10, 10, 10)."""
    end, _, _, _ = tracer.trace(seeds, reverse=True)
    lost = ~(end[:, 0] < 0.5)
    end[lost] = 10.0
    return end[:, 0].copy(), end[:, 1].copy(), end[:, 2].copy()


def make_rev_streamtrace_seeds(minx, maxx, miny, maxy, numpoints):
    """streamtrace.py:344-353: numpoints x numpoints lattice on the plane x = 3.9."""
    gy, gz = np.meshgrid(np.linspace(minx, maxx, num=numpoints), np.linspace(miny, maxy, num=numpoints))
    yz = np.stack((gy, gz), axis=-1).reshape(-1, 2)
    return np.hstack((np.full((yz.shape[0], 1), 3.9), yz))
