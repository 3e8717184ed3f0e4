#!/bin/bash
# round-2 evidence on L: plain bench line, ncu --set full capture of the assembly kernel, launch list of the bench command
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_L.json 2> gpurun_out/r2_bench_L.err; echo "bench rc=$?"
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_p1tet_ws -c 1 -o gpurun_out/r2_L_ws -f python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-aij > gpurun_out/r2_ncu_L.log 2>&1; echo "ncu full rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_L.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-aij > gpurun_out/r2_ncu_launches.log 2>&1; echo "ncu launches rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference_L.json 2> gpurun_out/r2_bench_reference_L.err; echo "ref rc=$?"
ls -la gpurun_out/r2_L_ws.ncu-rep gpurun_out/r2_launches_L.csv
