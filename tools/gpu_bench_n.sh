#!/bin/bash
# one bench line at N GPUs with the default options
N=${1:-8}; TAG=${2:-r2h}
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus $N --steps 10 --warmup 3 --no-aij > gpurun_out/${TAG}_bench_L_n${N}.json 2> gpurun_out/${TAG}_bench_L_n${N}.err; echo "bench rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_L_n${N}.json").read().strip().splitlines()[-1])
print("step",round(d["ms_per_step"],3),"kernel",round(d["roofline"]["kernel_ms"],3),"value",round(d["value"],1),"e2e",round(d["e2e"]["ms_per_step"],2),"spmv",round(d["spmv"]["ms"],3),"tfqmr",round(d["tfqmr"]["ms_per_iteration"],3),"ilu",d["tfqmr_ilu"].get("ms_per_iteration"),d["checksums"]["F_l2"],d["checksums"]["J_frobenius"])
PY
