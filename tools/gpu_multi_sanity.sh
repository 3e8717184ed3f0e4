#!/bin/bash
# N-GPU sanity of the final code: distributed parity, one bench line (default options)
N=${1:-4}; TAG=${2:-r2f}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 tests/multigpu_check.py > gpurun_out/${TAG}_multigpu_check_n$N.log 2>&1; echo "check rc=$?"
grep -E "^rank" gpurun_out/${TAG}_multigpu_check_n$N.log | cut -c1-200 | head -20
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29622 bench.py --gpus $N --steps 10 --warmup 3 --no-aij > gpurun_out/${TAG}_bench_L_n${N}.json 2> gpurun_out/${TAG}_bench_L_n${N}.err; echo "bench rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29623 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_ref_n${N}.json 2> gpurun_out/${TAG}_bench_ref_n${N}.err; echo "ref rc=$?"
python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_L_n${N}.json").read().strip().splitlines()[-1])
print("step",round(d["ms_per_step"],3),"kernel",round(d["roofline"]["kernel_ms"],3),"value",round(d["value"],1),"e2e",round(d["e2e"]["ms_per_step"],2),"spmv",round(d["spmv"]["ms"],3),"tfqmr",round(d["tfqmr"]["ms_per_iteration"],3),d["checksums"]["F_l2"],d["checksums"]["J_frobenius"])
print(open("gpurun_out/${TAG}_bench_ref_n${N}.json").read()[:300])
PY
