#!/bin/bash
# what the driver runs at round end, on one B200: GPU test suite, smoke, default bench line, reference arm
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2v_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2v_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2v_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r2v_smoke.log
timeout 900 python bench.py > gpurun_out/r2v_bench_L.json 2> gpurun_out/r2v_bench_L.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2v_bench_reference_L.json 2> gpurun_out/r2v_bench_reference_L.err; echo "ref rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/r2v_bench_L.json").read().strip().splitlines()[-1])
print("step", round(d["ms_per_step"],3), "kernel", round(d["roofline"]["kernel_ms"],3), "frac", round(d["roofline"]["frac"],4), "e2e", round(d["e2e"]["ms_per_step"],2), "spmv", round(d["spmv"]["ms"],3), "tfqmr", round(d["tfqmr"]["ms_per_iteration"],3), "ilu", d["tfqmr_ilu"])
PY
