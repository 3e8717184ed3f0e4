#!/bin/bash
# round-2 GPU check: parity tests, then short benches of the warp-specialised kernel vs the pipelined one
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -5 gpurun_out/r2c_pytest.log
timeout 300 python bench.py --workload M --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_M_ws.json 2> gpurun_out/r2c_bench_M_ws.err; echo "M ws rc=$?"
timeout 300 python bench.py --workload M --steps 10 --warmup 3 --no-cpu-baseline --ws 0 > gpurun_out/r2c_bench_M_pipe.json 2> gpurun_out/r2c_bench_M_pipe.err; echo "M pipe rc=$?"
timeout 600 python bench.py --workload L --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_bench_L_ws.json 2> gpurun_out/r2c_bench_L_ws.err; echo "L ws rc=$?"
python - <<'PY'
import json
for f in ["r2c_bench_M_ws","r2c_bench_M_pipe","r2c_bench_L_ws"]:
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["roofline"]["kernel_ms"], d["config"].get("kernel"), d["e2e"]["ms_per_step"], d["residual_only"]["ms"], d["jacobian_only"]["ms"])
    except Exception as e:
        print(f, "ERR", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
