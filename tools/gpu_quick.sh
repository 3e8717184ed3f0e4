#!/bin/bash
# quick GPU check of a kernel change: the kernel-variant / numbering / scale parity tests, then M and L benches
TAG=${1:-q}
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
for W in M L; do
  timeout 600 python bench.py --workload $W --steps 10 --warmup 3 --no-cpu-baseline --no-aij > gpurun_out/${TAG}_bench_$W.json 2> gpurun_out/${TAG}_bench_$W.err; echo "$W rc=$?"
done
python - <<PY
import json
for f in ["${TAG}_bench_M","${TAG}_bench_L"]:
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "step", round(d["ms_per_step"],3), "kernel", round(d["roofline"]["kernel_ms"],3), d["config"].get("kernel"), "e2e", round(d["e2e"]["ms_per_step"],2), "F", round(d["residual_only"]["ms"],3), "J", round(d["jacobian_only"]["ms"],3), "fp64", round(d["fp64"]["frac"],3))
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
