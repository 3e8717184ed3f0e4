#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -15 gpurun_out/r2d_pytest.log
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_p1tet_ws -c 1 -o gpurun_out/prof_r2d_M_ws -f python bench.py --workload M --steps 2 --warmup 1 --no-cpu-baseline --no-aij > gpurun_out/r2d_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/prof_r2d_M_ws.ncu-rep
