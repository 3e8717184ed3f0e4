#!/bin/bash
# plain timing, then ncu --set full of the two J+F launches of the row-owner kernel (vertex class, edge class)
TAG=${1:-r2g}
mkdir -p gpurun_out
timeout 200 python tools/prof_rowown.py 24 2; echo "plain rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_rowown -s 4 -c 2 -o gpurun_out/${TAG}_rowown -f python tools/prof_rowown.py 24 2 > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/${TAG}_rowown.ncu-rep
