"""One J+F, one F-only assembly of a P2-P1 duct through the row-owner kernel (for ncu captures)."""
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from stabilized_navier_stokes_flow_fenicsx_b200 import mesh as M
from stabilized_navier_stokes_flow_fenicsx_b200.assembler import NSAssembler

n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
vdeg = int(sys.argv[2]) if len(sys.argv) > 2 else 2
m = M.duct_mesh(n, 2 * n); sp = M.mixed_space(m, vdeg)
asm = NSAssembler(m.x, m.cells, sp.dofmap, vdeg=sp.vdeg)
asm.set_form(flavour=0 if vdeg == 2 else 2, nu=1 / 50, beta=0.2); asm.set_bcs(M.duct_bcs(sp))
asm.create_matrix(fetch=False)
x_dev, F_dev = asm.dev_alloc(8 * asm.n_cols), asm.dev_alloc(8 * asm.n_cols)
asm.h2d(x_dev, M.duct_state(sp))
for it in range(3):
    asm.jacobian_residual_dev(x_dev, True, F_dev); t = asm.last_kernel_ms()
    asm.jacobian_residual_dev(x_dev, False, F_dev); t2 = asm.last_kernel_ms()
print(asm.last_kernel_name(), m.n_cells, "cells", "JF", t, "ms", "F", t2, "ms", m.n_cells / t / 1e3, "Mcells/s")
