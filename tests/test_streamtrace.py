"""Streamline tracing: the integrator the GPU runs (csrc/trace_core.cuh, compiled here with g++ by a
test-only harness) against scipy.solve_ivp called exactly as NavierStokes/streamtrace.py:198-218 / :357-384 call it."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

from oracle.streamtrace_oracle import TraceOracle, host_tables
from stabilized_navier_stokes_flow_fenicsx_b200 import mesh as M
from stabilized_navier_stokes_flow_fenicsx_b200 import streamtrace as ST

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def swirl_field(x):
    """duct flow with a cross-stream swirl so that the streamlines cross many cells and bend"""
    y, z = x[:, 1], x[:, 2]
    prof = (1 - 4 * y * y) * (1 - 4 * z * z)
    u = np.empty_like(x)
    u[:, 0] = 1.5 * prof * (1 + 0.1 * np.sin(2 * np.pi * x[:, 0])) + 0.02
    u[:, 1] = -0.6 * z * prof
    u[:, 2] = 0.6 * y * prof
    return u


def seeds_plane(n, x0=0.2, r=0.35, seed=3):
    rng = np.random.default_rng(seed)
    yz = rng.uniform(-r, r, size=(n, 2))
    return np.hstack((np.full((n, 1), x0), yz))


@pytest.fixture(scope="module")
def host():
    out = os.path.join(ROOT, "tests", "host", "libtrace_host.so")
    src = os.path.join(ROOT, "tests", "host", "trace_host.cpp")
    hdr = os.path.join(ROOT, "stabilized_navier_stokes_flow_fenicsx_b200", "csrc", "trace_core.cuh")
    if not os.path.exists(out) or os.path.getmtime(out) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", "-x", "c++", src, "-o", out])
    return ctypes.CDLL(out)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def expected_status(sol, reverse):
    """nsgpu_trace_run's code for a solve_ivp result (forward events = (speed, position), reverse = (position, speed))"""
    if sol.status == 0:
        return ST.REACHED_T_END
    pos, spd = (sol.t_events[0], sol.t_events[1]) if reverse else (sol.t_events[1], sol.t_events[0])
    return ST.EVENT_POSITION if pos.size and pos[-1] == sol.t[-1] else ST.EVENT_SPEED


def host_run(lib, T, seeds, reverse, x_stop, speed_min=1e-6, t_end=20.0, max_step=0.125, rtol=1e-3, atol=1e-6):
    n = seeds.shape[0]
    end = np.empty((n, 3)); status = np.empty(n, np.int32); tf = np.empty(n); ns = np.empty(n, np.int32); nf = np.empty(n, np.int32)
    lib.host_trace_run(_p(T["cmap"]), _p(T["cvel"]), _p(T["bin_ptr"]), _p(T["bin_cells"]), _p(T["lo"]), _p(T["inv_h"]), _p(T["nb"]),
                       ctypes.c_double(T["tol"]), ctypes.c_int64(n), _p(np.ascontiguousarray(seeds)), ctypes.c_int(reverse), ctypes.c_double(x_stop),
                       ctypes.c_double(speed_min), ctypes.c_double(t_end), ctypes.c_double(max_step), ctypes.c_double(rtol), ctypes.c_double(atol),
                       ctypes.c_int64(1000000), _p(end), _p(status), _p(tf), _p(ns), _p(nf))
    return end, status, tf, ns, nf


@pytest.fixture(scope="module")
def duct():
    m = M.duct_mesh(6, 24)
    u = swirl_field(m.x)
    return m, u, TraceOracle(m.x, m.cells, u), host_tables(m.x, m.cells, u)


def test_velfunc_matches_the_oracle_inside_and_outside(host, duct):
    m, u, orc, T = duct
    rng = np.random.default_rng(0)
    pts = np.column_stack((rng.uniform(-0.2, 4.2, 400), rng.uniform(-0.6, 0.6, 400), rng.uniform(-0.6, 0.6, 400)))
    pts[:20] = m.x[rng.integers(0, m.x.shape[0], 20)]            # mesh vertices (shared by many cells)
    vel = np.empty_like(pts); cell = np.empty(400, np.int32)
    host.host_trace_velocity(_p(T["cmap"]), _p(T["cvel"]), _p(T["bin_ptr"]), _p(T["bin_cells"]), _p(T["lo"]), _p(T["inv_h"]), _p(T["nb"]),
                             ctypes.c_double(T["tol"]), ctypes.c_int64(400), _p(pts), _p(vel), _p(cell))
    ref = np.array([orc.velfunc(0.0, p) for p in pts])
    inside = np.array([orc.locate(p)[0] >= 0 for p in pts])
    assert inside.sum() > 100 and (~inside).sum() > 50
    assert np.array_equal(cell >= 0, inside)
    assert np.all(vel[~inside] == 0.0)
    assert np.abs(vel - ref).max() < 1e-13


def test_forward_trace_follows_scipy_step_for_step(host, duct):
    m, u, orc, T = duct
    seeds = seeds_plane(12)
    end, status, tf, ns, nf = host_run(host, T, seeds, 0, 3.7)
    for i, s in enumerate(seeds):
        sol = orc.forward(s)
        assert status[i] == expected_status(sol, False)
        assert ns[i] == sol.t.size - 1 and nf[i] == sol.nfev       # same accepted steps and right-hand-side calls
        assert abs(tf[i] - sol.t[-1]) < 1e-10
        assert np.abs(end[i] - sol.y[:, -1]).max() < 1e-10
        if status[i] == ST.EVENT_POSITION:
            assert abs(end[i, 0] - 3.7) < 1e-12
    assert (status == ST.EVENT_POSITION).sum() >= 8


def test_reverse_trace_and_speed_event(host, duct):
    m, u, orc, T = duct
    seeds = np.hstack((np.full((6, 1), 3.9), seeds_plane(6, r=0.3)[:, 1:]))
    end, status, tf, ns, nf = host_run(host, T, seeds, 1, 0.13)
    for i, s in enumerate(seeds):
        sol = orc.reverse(s)
        assert status[i] == expected_status(sol, True)
        assert ns[i] == sol.t.size - 1
        assert np.abs(end[i] - sol.y[:, -1]).max() < 1e-10
    # a seed whose streamline leaves through the wall region / a seed outside the mesh: the speed event ends the trace
    out = np.array([[0.5, 0.7, 0.0], [4.5, 0.0, 0.0]])
    end, status, tf, ns, nf = host_run(host, T, out, 0, 3.7)
    for i, s in enumerate(out):
        sol = orc.forward(s)
        assert np.abs(end[i] - sol.y[:, -1]).max() < 1e-12
        assert abs(tf[i] - sol.t[-1]) < 1e-12
    # t_end reached before any event (short horizon)
    end, status, tf, ns, nf = host_run(host, T, seeds_plane(3), 0, 3.7, t_end=0.5)
    assert np.all(status == ST.REACHED_T_END) and np.all(tf == 0.5)
    for i, s in enumerate(seeds_plane(3)):
        sol = orc.forward(s, t_end=0.5)
        assert sol.status == 0 and np.abs(end[i] - sol.y[:, -1]).max() < 1e-10


def test_seed_lattice_of_the_reverse_trace():
    s = ST.make_rev_streamtrace_seeds(-0.2, 0.3, -0.1, 0.1, 5)
    assert s.shape == (25, 3) and np.all(s[:, 0] == 3.9)
    assert np.allclose(s[:5, 1], np.linspace(-0.2, 0.3, 5)) and np.allclose(s[::5, 2], np.linspace(-0.1, 0.1, 5))


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_gpu_velfunc_and_traces_match_scipy(duct):
    m, u, orc, T = duct
    tr = ST.StreamTracer(m.x, m.cells, u)
    rng = np.random.default_rng(1)
    pts = np.column_stack((rng.uniform(-0.2, 4.2, 2000), rng.uniform(-0.6, 0.6, 2000), rng.uniform(-0.6, 0.6, 2000)))
    vel, cell = tr.velfunc(pts, return_cells=True)
    ref = np.array([orc.velfunc(0.0, p) for p in pts[:300]])
    assert np.abs(vel[:300] - ref).max() < 1e-13
    assert np.all(vel[cell < 0] == 0.0) and (cell < 0).sum() > 100
    seeds = seeds_plane(16)
    end, status, tf, ns = tr.trace(seeds)
    for i, s in enumerate(seeds):
        sol = orc.forward(s)
        assert status[i] == expected_status(sol, False) and ns[i] == sol.t.size - 1
        assert np.abs(end[i] - sol.y[:, -1]).max() < 1e-9 and abs(tf[i] - sol.t[-1]) < 1e-9
    rseeds = np.hstack((np.full((8, 1), 3.9), seeds_plane(8, r=0.3)[:, 1:]))
    end, status, tf, ns = tr.trace(rseeds, reverse=True)
    for i, s in enumerate(rseeds):
        sol = orc.reverse(s)
        assert np.abs(end[i] - sol.y[:, -1]).max() < 1e-9
    assert tr.lib.nsgpu_last_kernel_name(tr.ctx) == b"trace_rk45"
    tr.close()


@pytest.mark.gpu
def test_gpu_forty_thousand_seeds_round_trip():
    """The reference's seed count on a finer duct: forward to x = 3.7, then the end points traced back upstream must return to
    their seeds' neighbourhood (size-independent property), and the filters of run_streamtrace / run_reverse_streamtrace hold."""
    m = M.duct_mesh(16, 64)
    u = swirl_field(m.x)
    tr = ST.StreamTracer(m.x, m.cells, u)
    seeds = seeds_plane(40000, x0=0.3, r=0.3, seed=5)
    end, status, tf, ns = tr.trace(seeds)
    assert np.all(status == ST.EVENT_POSITION) and np.abs(end[:, 0] - 3.7).max() < 1e-12
    back, bstatus, _, _ = tr.trace(end, reverse=True, x_stop=0.3)
    assert np.all(bstatus == ST.EVENT_POSITION)
    assert np.abs(back - seeds).max() < 5e-2 and np.median(np.abs(back - seeds).max(axis=1)) < 1e-2
    px, py, pz = ST.run_streamtrace(seeds[:1000], tr)
    assert px.shape == (1000, 1) and np.all(px > 0.5)
    far = np.vstack((ST.make_rev_streamtrace_seeds(-0.2, 0.2, -0.2, 0.2, 10), [[3.9, 0.9, 0.9]]))
    rx, ry, rz = ST.run_reverse_streamtrace(far, tr)
    assert rx.shape == (101,) and rx[-1] == 10.0 and ry[-1] == 10.0 and np.all(rx[:-1] < 0.5)
    tr.close()
