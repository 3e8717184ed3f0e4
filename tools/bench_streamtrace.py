"""Streamline tracing throughput: the reference's 40 000 seeds (streamtrace.py:666-668 numpoints = 200) on a duct mesh, GPU
(all seeds in one nsgpu_trace_run call, end to end from host seeds to host end points) vs the reference's own per-seed
scipy.solve_ivp loop on a small sample of the same seeds (oracle right-hand side).  Prints one JSON line."""
import json
import sys
import time

import numpy as np

sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
from stabilized_navier_stokes_flow_fenicsx_b200 import mesh as M
from stabilized_navier_stokes_flow_fenicsx_b200 import streamtrace as ST


def field(x):
    y, z = x[:, 1], x[:, 2]
    prof = (1 - 4 * y * y) * (1 - 4 * z * z)
    return np.column_stack((1.5 * prof * (1 + 0.1 * np.sin(2 * np.pi * x[:, 0])) + 0.02, -0.6 * z * prof, 0.6 * y * prof))


def main():
    n_cross, n_long = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (50, 200)
    n_seeds = int(sys.argv[3]) if len(sys.argv) > 3 else 40000
    m = M.duct_mesh(n_cross, n_long)
    u = field(m.x)
    t0 = time.perf_counter()
    tr = ST.StreamTracer(m.x, m.cells, u)
    t_setup = time.perf_counter() - t0
    rng = np.random.default_rng(5)
    seeds = np.hstack((np.full((n_seeds, 1), 0.3), rng.uniform(-0.3, 0.3, size=(n_seeds, 2))))
    tr.trace(seeds[:256])
    best, kern = 1e30, 0.0
    for _ in range(3):
        t0 = time.perf_counter()
        end, status, tf, ns = tr.trace(seeds)
        dt = time.perf_counter() - t0
        if dt < best:
            best, kern = dt, tr.last_kernel_ms()
    out = {"what": "streamtrace forward, RK45 rtol 1e-3 max_step 0.125, events at x = 3.7", "cells": int(m.cells.shape[0]), "seeds": n_seeds,
           "gpu_ms_e2e": best * 1e3, "gpu_ms_kernel": kern, "gpu_seeds_per_s": n_seeds / best, "setup_s": t_setup,
           "accepted_steps_total": int(ns.sum()), "reached_outlet": int((status == 1).sum())}
    if m.cells.shape[0] <= 200000:
        from oracle.streamtrace_oracle import TraceOracle
        orc = TraceOracle(m.x, m.cells, u)
        k = 4
        t0 = time.perf_counter()
        err = 0.0
        for i in range(k):
            sol = orc.forward(seeds[i])
            err = max(err, float(np.abs(sol.y[:, -1] - end[i]).max()))
        dt = (time.perf_counter() - t0) / k
        out.update({"scipy_s_per_seed": dt, "scipy_seeds_per_s": 1.0 / dt, "max_abs_diff_vs_scipy": err,
                    "cpu_note": "scipy.solve_ivp per seed as the reference calls it; right-hand side = NumPy brute-force locator (restated), 1 core"})
    print(json.dumps(out))


if __name__ == "__main__":
    main()
