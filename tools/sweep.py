#!/usr/bin/env python
"""tools/sweep.py -- time several option sets of the factorised P1-P1 kernel in ONE process (one mesh / pattern build).

  python tools/sweep.py --workload M --variants "ws=0;ws=1;ws=1,debug=16;threads=64"

For every variant: 3 warm-up + --steps timed fused J+F assemblies (kernel time from the library's CUDA events), and -- unless
a `debug` timing switch is on -- a comparison of the assembled values and residual with the first variant (max |diff| / max |ref|).
One JSON line per variant.  Development tool; bench.py is the contract benchmark.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

WORKLOADS = {"S": (10, 40), "M": (50, 200), "L": (128, 512), "M2": (64, 256)}
DEFAULTS = {"kernel": 0, "threads": 128, "lanes": 1, "ws": 0, "debug": 0, "persistent": 1, "pipe": 1}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--workload", default="M", choices=sorted(WORKLOADS))
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--variants", default="ws=0;ws=1")
    ap.add_argument("--no-check", action="store_true")
    args = ap.parse_args()

    from stabilized_navier_stokes_flow_fenicsx_b200 import mesh as M
    from stabilized_navier_stokes_flow_fenicsx_b200 import distributed as D
    from stabilized_navier_stokes_flow_fenicsx_b200.assembler import NSAssembler

    n_cross, n_long = WORKLOADS[args.workload]
    part = D.duct_partition(n_cross, n_long, 0, 1)
    asm = NSAssembler(part.x, part.cells, part.dofmap, vdeg=1, n_dofs_owned=part.n_owned, n_dofs_ghost=part.n_ghost,
                      n_cells_owned=part.n_cells_owned, device=0)
    asm.set_form(flavour=0, nu=0.1, Ci=36.0)
    asm.set_bcs(part.bcs)
    asm.create_matrix(fetch=False)
    nbytes = 8 * asm.n_cols
    x_dev, F_dev = asm.dev_alloc(nbytes), asm.dev_alloc(nbytes)
    w = np.zeros(asm.n_cols)
    w[: asm.n_dofs] = part.w
    asm.h2d(x_dev, w)
    nc = 6 * n_cross * n_cross * n_long
    check = not args.no_check and args.workload != "L"
    ref_vals = ref_F = None
    Fh = np.zeros(asm.n_cols)
    for spec in args.variants.split(";"):
        opts = dict(DEFAULTS)
        for kv in filter(None, spec.split(",")):
            k, v = kv.split("=")
            opts[k.strip()] = int(v)
        for k, v in opts.items():
            asm.set_option(k, v)
        out = {"variant": spec, "workload": args.workload}
        try:
            for _ in range(3):
                asm.jacobian_residual_dev(x_dev, True, F_dev)
            ms = []
            for _ in range(args.steps):
                asm.jacobian_residual_dev(x_dev, True, F_dev)
                ms.append(asm.last_kernel_ms())
            out.update(kernel_ms=float(np.mean(ms)), kernel_ms_min=float(np.min(ms)), Mcells_s=nc / (np.mean(ms) * 1e-3) / 1e6)
            fms = []
            for _ in range(args.steps):
                asm.jacobian_residual_dev(x_dev, False, F_dev)
                fms.append(asm.last_kernel_ms())
            out.update(residual_ms=float(np.mean(fms)))
            if check and not opts["debug"] & 15:
                asm.jacobian_residual_dev(x_dev, True, F_dev)
                vals = asm.get_values()
                asm.d2h(Fh, F_dev)
                if ref_vals is None:
                    ref_vals, ref_F = vals, Fh.copy()
                else:
                    out["vals_rel_diff"] = float(np.abs(vals - ref_vals).max() / np.abs(ref_vals).max())
                    out["F_rel_diff"] = float(np.abs(Fh - ref_F)[: asm.n_owned].max() / np.abs(ref_F).max())
        except Exception as e:   # keep sweeping: a failing variant must not hide the others
            out["error"] = str(e)[:300]
        print(json.dumps(out), flush=True)
    asm.close()


if __name__ == "__main__":
    main()
