#!/bin/bash
# row-owner kernel check: parity tests, then the other configurations
TAG=${1:-r2f}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_newton_gpu.py tests/test_reference_pins.py -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${TAG}_pytest.log
tail -6 gpurun_out/${TAG}_pytest.log
timeout 300 python tools/bench_configs.py > gpurun_out/${TAG}_other_configs.jsonl 2> gpurun_out/${TAG}_other_configs.err; echo "configs rc=$?"
python - <<PY
import json
for l in open("gpurun_out/${TAG}_other_configs.jsonl"):
    d=json.loads(l); print(d["case"][:70].ljust(70), d["kernel"], "jf", round(d["jf_ms"],3), "f", round(d["f_ms"],3), round(d["jf_Mcells_s"],1))
PY
tail -3 gpurun_out/${TAG}_other_configs.err
