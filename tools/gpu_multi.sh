#!/bin/bash
# multi-GPU check on N GPUs of one box: distributed parity (tests/multigpu_check.py) and the scaling bench line
N=${1:-2}; TAG=${2:-mg}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 tests/multigpu_check.py > gpurun_out/${TAG}_multigpu_check_n$N.log 2>&1; echo "check rc=$?"
tail -12 gpurun_out/${TAG}_multigpu_check_n$N.log
for OV in 1 0; do
  NSGPU_OVERLAP=$OV timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 10 --warmup 3 --no-aij > gpurun_out/${TAG}_bench_L_n${N}_ov$OV.json 2> gpurun_out/${TAG}_bench_L_n${N}_ov$OV.err; echo "bench ov=$OV rc=$?"
done
python - <<PY
import json
for ov in (1,0):
    f=f"gpurun_out/${TAG}_bench_L_n${N}_ov{ov}.json"
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print("overlap",ov,"step",round(d["ms_per_step"],3),"kernel",round(d["roofline"]["kernel_ms"],3),"value",round(d["value"],1),"e2e",round(d["e2e"]["ms_per_step"],2),"spmv",round(d["spmv"]["ms"],3),d["checksums"])
    except Exception as e:
        print(f,"ERR",e); print(open(f.replace(".json",".err")).read()[-2000:])
PY
