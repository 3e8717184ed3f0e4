#!/bin/bash
# ncu --set full capture (with source) of the assembly kernel on workload $1 (default M), tag $2
W=${1:-M}; TAG=${2:-prof}
mkdir -p gpurun_out
timeout 900 ncu --set full --import-source on --clock-control none -k regex:k_p1tet_ws -c 1 -o gpurun_out/${TAG}_${W}_ws -f python bench.py --workload $W --steps 2 --warmup 1 --no-cpu-baseline --no-aij > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/${TAG}_${W}_ws.ncu-rep
