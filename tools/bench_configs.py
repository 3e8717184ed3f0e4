#!/usr/bin/env python
"""Throughput of the other BASELINE configurations (generic kernel): UGN triangles (config 2, lid-driven cavity),
P2-P1 Taylor-Hood tets (config 4), the Stokes operators -- J+F device time per call, CUDA events, median of 10."""
import json
import sys
import os
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stabilized_navier_stokes_flow_fenicsx_b200 import mesh as M
from stabilized_navier_stokes_flow_fenicsx_b200.assembler import NSAssembler


def run(name, m, space, w, bcs, rowown=2, **form):
    sp = space
    asm = NSAssembler(m.x, m.cells, sp.dofmap, vdeg=sp.vdeg)
    asm.set_form(**form); asm.set_bcs(bcs)
    asm.set_option("rowown", rowown)
    asm.create_matrix(fetch=False)
    x_dev, F_dev, y_dev = (asm.dev_alloc(8 * asm.n_cols) for _ in range(3))
    asm.h2d(x_dev, w)
    jf, f, sp_ms = [], [], []
    for it in range(13):
        asm.jacobian_residual_dev(x_dev, True, F_dev); t = asm.last_kernel_ms()
        asm.jacobian_residual_dev(x_dev, False, F_dev); t2 = asm.last_kernel_ms()
        asm.spmv_dev(x_dev, y_dev); t3 = asm.last_kernel_ms()
        if it >= 3:
            jf.append(t); f.append(t2); sp_ms.append(t3)
    r = {"case": name, "kernel": asm.last_kernel_name(), "cells": m.n_cells, "dofs": sp.n_dofs, "nnz": asm.nnz, "jf_ms": float(np.median(jf)), "f_ms": float(np.median(f)),
         "spmv_ms": float(np.median(sp_ms)), "j_ms": None, "jf_Mcells_s": m.n_cells / np.median(jf) / 1e3, "spmv_GBs": (12 * asm.nnz + 24 * sp.n_dofs) / np.median(sp_ms) / 1e6}
    print(json.dumps(r), flush=True)
    asm.close()


if __name__ == "__main__":
    m = M.create_rectangle_tris(64, 64); sp = M.mixed_space(m, 1)
    run("config2 lid-driven cavity N=64, UGN P1-P1 triangles", m, sp, M.cavity_state(sp), M.cavity_bcs(sp), flavour=1, nu=1 / 100)
    m = M.create_rectangle_tris(1024, 1024); sp = M.mixed_space(m, 1)
    run("cavity N=1024, UGN P1-P1 triangles", m, sp, M.cavity_state(sp), M.cavity_bcs(sp), flavour=1, nu=1 / 100)
    run("cavity N=1024, UGN P1-P1 triangles, cooperative kernel with atomics", m, sp, M.cavity_state(sp), M.cavity_bcs(sp), rowown=0, flavour=1, nu=1 / 100)
    m = M.duct_mesh(40, 80); sp = M.mixed_space(m, 2)
    run("config4-like duct 40x40x80, G-metric P2-P1 tets", m, sp, M.duct_state(sp), M.duct_bcs(sp), flavour=0, nu=1 / 50)
    run("config4-like duct 40x40x80, G-metric P2-P1 tets, cooperative kernel with atomics", m, sp, M.duct_state(sp), M.duct_bcs(sp), rowown=0, flavour=0, nu=1 / 50)
    run("duct 40x40x80, Stokes P2-P1 (DuctStokesFlow form)", m, sp, M.duct_state(sp), M.duct_bcs(sp), flavour=2, nu=1.0, alpha=1.0, sp=-1.0, beta=0.0)
    m = M.duct_mesh(50, 200); sp = M.mixed_space(m, 1)
    run("duct M, Stokes P1-P1 + PSPG (channel initial guess)", m, sp, M.duct_state(sp), M.duct_bcs(sp), flavour=2, nu=1.0, alpha=1.0, sp=1.0, beta=0.2)
    run("duct M, G-metric P1-P1 (row-owner kernel)", m, sp, M.duct_state(sp), M.duct_bcs(sp), flavour=0, nu=0.1)
