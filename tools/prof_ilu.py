"""A few ILU-preconditioned TFQMR iterations on the duct (for ncu captures of the sweep kernels)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stabilized_navier_stokes_flow_fenicsx_b200 import mesh as M
from stabilized_navier_stokes_flow_fenicsx_b200.assembler import NSAssembler

nc, nl = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (64, 256)
m = M.duct_mesh(nc, nl); sp = M.mixed_space(m, 1)
asm = NSAssembler(m.x, m.cells, sp.dofmap, vdeg=1)
asm.set_form(flavour=0, nu=0.1); asm.set_bcs(M.duct_bcs(sp))
asm.create_matrix(fetch=False)
x_dev, F_dev, y_dev = (asm.dev_alloc(8 * asm.n_cols) for _ in range(3))
asm.h2d(x_dev, M.duct_state(sp))
asm.jacobian_residual_dev(x_dev, True, F_dev)
import time
for its in (1, 3, 13):
    asm.sync(); t0 = time.perf_counter()
    info = asm.tfqmr_dev(F_dev, y_dev, rtol=0.0, max_it=its, pc=5)
    asm.sync(); print(its, "iterations", 1e3 * (time.perf_counter() - t0), "ms", info)
print(m.n_cells, "cells", asm.ilu_colours())
