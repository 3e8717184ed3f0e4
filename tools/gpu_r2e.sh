#!/bin/bash
# round-2 session-4 check: full GPU test suite, the other configurations, a short L bench
TAG=${1:-r2e}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
timeout 300 python tools/bench_configs.py > gpurun_out/${TAG}_other_configs.jsonl 2> gpurun_out/${TAG}_other_configs.err; echo "configs rc=$?"
cut -c1-400 gpurun_out/${TAG}_other_configs.jsonl
timeout 400 python bench.py --workload L --steps 10 --warmup 3 --no-cpu-baseline --no-aij > gpurun_out/${TAG}_bench_L.json 2> gpurun_out/${TAG}_bench_L.err; echo "L rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_L.json").read().strip().splitlines()[-1])
print("step", round(d["ms_per_step"],3), "kernel", round(d["roofline"]["kernel_ms"],3), d["config"].get("kernel"), "tfqmr", d["tfqmr"]["ms_per_iteration"], "ilu", d.get("tfqmr_ilu"))
PY
