#!/bin/bash
# final round-2 evidence on one B200: full test suite, the default bench line on L, the reference arm, the other configurations,
# ncu captures (SpMV on L, row-owner P2-P1 kernels) and the launch list of the bench command
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2z_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r2z_pytest.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2z_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2z_smoke.log
timeout 900 python bench.py > gpurun_out/r2z_bench_L.json 2> gpurun_out/r2z_bench_L.err; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2z_bench_reference_L.json 2> gpurun_out/r2z_bench_reference_L.err; echo "ref rc=$?"
timeout 300 python tools/bench_configs.py > gpurun_out/r2z_other_configs.jsonl 2> gpurun_out/r2z_other_configs.err; echo "configs rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_spmv_block4 -s 3 -c 1 -o gpurun_out/r2z_L_spmv -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-aij --no-extras > gpurun_out/r2z_ncu_spmv.log 2>&1; echo "ncu spmv rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2z_launches_L.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-aij --no-extras > gpurun_out/r2z_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_rowown -s 4 -c 2 -o gpurun_out/r2z_rowown -f python tools/prof_rowown.py 32 2 > gpurun_out/r2z_ncu_rowown.log 2>&1; echo "ncu rowown rc=$?"
ls -la gpurun_out/r2z_*
