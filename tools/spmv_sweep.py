#!/usr/bin/env python
"""tools/spmv_sweep.py -- vertex-blocked SpMV time on a duct for the resident-CTA settings of the kernel (option spmv_blocks)."""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stabilized_navier_stokes_flow_fenicsx_b200 import distributed as D
from stabilized_navier_stokes_flow_fenicsx_b200.assembler import NSAssembler

wl = {"M": (50, 200), "L": (128, 512)}[sys.argv[1] if len(sys.argv) > 1 else "L"]
part = D.duct_partition(wl[0], wl[1], 0, 1)
asm = NSAssembler(part.x, part.cells, part.dofmap, vdeg=1, n_dofs_owned=part.n_owned, n_dofs_ghost=part.n_ghost, n_cells_owned=part.n_cells_owned)
asm.set_form(flavour=0, nu=0.1); asm.set_bcs(part.bcs); asm.create_matrix(fetch=False)
x_dev, y_dev, F_dev = (asm.dev_alloc(8 * asm.n_cols) for _ in range(3))
w = np.zeros(asm.n_cols); w[: asm.n_dofs] = part.w
asm.h2d(x_dev, w)
asm.jacobian_residual_dev(x_dev, True, F_dev)
ref = None
yh = np.zeros(asm.n_cols)
for mb in (4, 5, 6):
    asm.set_option("spmv_blocks", mb)
    ms = []
    for it in range(13):
        asm.spmv_dev(x_dev, y_dev); t = asm.last_kernel_ms()
        if it >= 3: ms.append(t)
    asm.d2h(yh, y_dev)
    cs = float(np.abs(yh[: asm.n_owned]).sum())
    ref = cs if ref is None else ref
    print(json.dumps({"spmv_blocks": mb, "ms": float(np.mean(ms)), "same_y": cs == ref}), flush=True)
asm.close()
