#!/bin/bash
# guarded check of an experimental kernel: short timeouts so that a deadlock costs a minute, not twenty
TAG=${1:-g}
mkdir -p gpurun_out
timeout 120 python bench.py --workload M --steps 10 --warmup 3 --no-cpu-baseline --no-aij > gpurun_out/${TAG}_bench_M.json 2> gpurun_out/${TAG}_bench_M.err; echo "M rc=$?"
if [ -s gpurun_out/${TAG}_bench_M.json ]; then
  timeout 200 python bench.py --workload L --steps 10 --warmup 3 --no-cpu-baseline --no-aij > gpurun_out/${TAG}_bench_L.json 2> gpurun_out/${TAG}_bench_L.err; echo "L rc=$?"
  timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -x -q 2>&1 | tail -3
fi
python - <<PY
import json
for f in ["${TAG}_bench_M","${TAG}_bench_L"]:
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "step", round(d["ms_per_step"],3), "kernel", round(d["roofline"]["kernel_ms"],3), d["config"].get("kernel"), "e2e", round(d["e2e"]["ms_per_step"],2), "F", round(d["residual_only"]["ms"],3), "J", round(d["jacobian_only"]["ms"],3))
    except Exception as e:
        print(f, "ERR", e)
PY
