#!/bin/bash
# DDP runs of bench.py on N GPUs of one box: $1 = N, $2 = workload, $3 = mode (infer|train), extra args after
N=$1; W=$2; M=$3; shift 3
export NCCL_DEBUG=${NCCL_DEBUG:-WARN}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus $N --workload $W --mode $M "$@" > gpurun_out/r02_ddp_${W}_${M}_n$N.json 2> gpurun_out/r02_ddp_${W}_${M}_n$N.err
echo "N=$N $W $M rc=$?"
tail -c 600 gpurun_out/r02_ddp_${W}_${M}_n$N.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02_ddp_${W}_${M}_n$N.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", (d.get("e2e") or {}).get("value"))
    print("train", d.get("train") or d.get("train_step"))
except Exception as e:
    print("no json", e)
PY
