#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/prof_ilu.py 64 256; echo "plain rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_ilu_sweep -s 30 -c 4 -o gpurun_out/r2w_ilu -f python tools/prof_ilu.py 64 256 > gpurun_out/r2w_ncu.log 2>&1; echo "ncu rc=$?"
