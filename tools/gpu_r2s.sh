#!/bin/bash
# Krylov / ILU check: tests, then the L bench line (TFQMR and ILU timings)
TAG=${1:-r2s}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ilu_gpu.py tests/test_krylov_gpu.py tests/test_newton_gpu.py -m gpu -x -q 2>&1 | tail -2
timeout 400 python bench.py --workload L --steps 10 --warmup 3 --no-cpu-baseline --no-aij > gpurun_out/${TAG}_bench_L.json 2> gpurun_out/${TAG}_bench_L.err; echo "L rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_L.json").read().strip().splitlines()[-1])
print("step", round(d["ms_per_step"],3), "spmv", round(d["spmv"]["ms"],3), "tfqmr", d["tfqmr"]["ms_per_iteration"], d["tfqmr"]["per_solve_overhead_ms"], "ilu", d.get("tfqmr_ilu"))
print(d.get("other_paths"))
PY
