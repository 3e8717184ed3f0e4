#!/bin/bash
# One command that turns "parity unpinned" into "pinned" on the first machine that has BOTH a B200 and FEniCSx:
#   conda env create -f <reference>/environment.yml && conda activate fenicsx-env     (the reference's own environment)
#   tools/run_dolfinx_parity.sh [n_cells_per_edge]
# It builds libnsgpu.so, runs the reference's own assembly sequence (fem.form, create_matrix, assemble_matrix(bcs),
# assemble_vector, apply_lifting(-1), set_bc(-1); NavierStokes/NavierStokesChannelFlow.py:40-75, 271-272) with dolfinx on a
# tetrahedral box, hands the SAME arrays (mesh.geometry.x, geometry.dofmap, W.dofmap.list, dirichletbc dofs/values) to the
# C ABI and compares: CSR pattern bit-equal, entries and residual to 1e-12 (tools/compare_with_dolfinx.py); then the
# pytest wrapper of the same comparison.  Exit code 0 = pinned.
set -euo pipefail
cd "$(dirname "$0")/.."
python - <<'PY'
import importlib, sys
missing = [m for m in ("dolfinx", "ufl", "basix", "petsc4py", "mpi4py") if importlib.util.find_spec(m) is None]
if missing:
    sys.exit("FEniCSx is not importable here (missing: %s): activate the reference's environment.yml first" % ", ".join(missing))
PY
python -c "import __graft_entry__ as g; g.build()"
N=${1:-8}
mkdir -p gpurun_out
python tools/compare_with_dolfinx.py --n "$N" "$N" "$((2 * N))" | tee gpurun_out/dolfinx_parity.log
python tools/compare_with_dolfinx.py --n "$N" "$N" "$((2 * N))" --p2 | tee -a gpurun_out/dolfinx_parity.log
python -m pytest tests/test_dolfinx_parity.py -q
echo "dolfinx parity: pinned (log in gpurun_out/dolfinx_parity.log; copy it to profiles/ and drop the 'unpinned' note in DESIGN.md section 2)"
