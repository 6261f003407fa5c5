#!/bin/bash
# where the multi-rank training step's time goes: TIGER_BENCH_DDP_EXP=1 re-times the loop without a collective, with one
# bucket after the backward pass, and with the three overlapped slices ($1 = N)
N=${1:-2}; shift
export NCCL_DEBUG=${NCCL_DEBUG:-WARN} TIGER_BENCH_DDP_EXP=1
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
  bench.py --gpus $N --workload scaled --mode train --steps 60 --warmup 10 "$@" \
  > gpurun_out/ddp_exp_n$N.json 2> gpurun_out/ddp_exp_n$N.err
echo rc=$?; grep "ddp-exp" gpurun_out/ddp_exp_n$N.err; tail -c 400 gpurun_out/ddp_exp_n$N.json
