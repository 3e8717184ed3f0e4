#!/usr/bin/env python
"""tools/e2e_chunks.py -- end-to-end (pinned host vectors) J+F time of the streamed path for several chunk counts."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from stabilized_navier_stokes_flow_fenicsx_b200 import distributed as D
from stabilized_navier_stokes_flow_fenicsx_b200.assembler import NSAssembler

wl = {"M": (50, 200), "L": (128, 512)}[sys.argv[1] if len(sys.argv) > 1 else "L"]
part = D.duct_partition(wl[0], wl[1], 0, 1)
asm = NSAssembler(part.x, part.cells, part.dofmap, vdeg=1, n_dofs_owned=part.n_owned, n_dofs_ghost=part.n_ghost, n_cells_owned=part.n_cells_owned)
asm.set_form(flavour=0, nu=0.1); asm.set_bcs(part.bcs); asm.create_matrix(fetch=False)
xh, Fh = asm.pinned_empty(asm.n_dofs), asm.pinned_empty(asm.n_dofs)
xh[:] = part.w
ref = None
for k in [int(a) for a in (sys.argv[2:] or ["8", "16", "32", "0"])]:
    if k == 0:
        asm.set_option("stream_host", 0)
    else:
        asm.set_option("stream_host", 1); asm.set_option("stream_chunks", k)
    for _ in range(2):
        asm.jacobian_residual(xh, F_out=Fh, fetch_vals=False)
    t0 = time.perf_counter()
    for _ in range(5):
        asm.jacobian_residual(xh, F_out=Fh, fetch_vals=False)
    ms = 1e3 * (time.perf_counter() - t0) / 5
    cs = float(np.abs(Fh).sum())
    ref = cs if ref is None else ref
    print(json.dumps({"chunks": k, "e2e_ms": ms, "kernel": asm.last_kernel_name(), "same_F": cs == ref}), flush=True)
asm.close()
