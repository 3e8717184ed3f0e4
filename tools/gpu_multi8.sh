#!/bin/bash
# 8-GPU evidence: distributed parity log, then the scaling bench with serial / overlapped exchanges
N=${1:-8}; TAG=${2:-mg8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tests/multigpu_check.py > gpurun_out/${TAG}_multigpu_check_n$N.log 2>&1; echo "check rc=$?"
grep -c "^rank" gpurun_out/${TAG}_multigpu_check_n$N.log
for CFG in "0 0" "1 0" "1 4"; do
  set -- $CFG
  NSGPU_OVERLAP=$1 NSGPU_SM_RESERVE=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 20 --warmup 3 --no-aij > gpurun_out/${TAG}_bench_L_n${N}_ov$1_r$2.json 2> gpurun_out/${TAG}_bench_L_n${N}_ov$1_r$2.err; echo "bench ov=$1 reserve=$2 rc=$?"
done
python - <<PY
import json,glob
for f in sorted(glob.glob("gpurun_out/${TAG}_bench_L_n${N}_ov*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("_n")[1], "step",round(d["ms_per_step"],3),"kernel",round(d["roofline"]["kernel_ms"],3),"value",round(d["value"],1),"e2e",round(d["e2e"]["ms_per_step"],2),"spmv",round(d["spmv"]["ms"],3),d["checksums"]["F_l2"],d["checksums"]["J_frobenius"])
    except Exception as e:
        print(f,"ERR",e); print(open(f.replace(".json",".err")).read()[-1500:])
PY
