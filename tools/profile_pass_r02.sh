#!/bin/bash
# Round-2 measurement pass (one B200 through gpurun).  Every ncu command runs only after the same command exited 0
# without ncu; only text artefacts travel back (gpurun_out is capped at 64 MiB).
TAG=${1:-r02}
O=gpurun_out
python __graft_entry__.py smoke > $O/${TAG}_smoke.log 2>&1; echo smoke_rc=$?
python -m pytest tests -m gpu -x -q > $O/${TAG}_pytest_gpu.log 2>&1; echo pytest_rc=$?; tail -2 $O/${TAG}_pytest_gpu.log
for w in reddit wikipedia mooc lastfm scaled; do
  python bench.py --workload $w > $O/${TAG}_bench_$w.json 2> $O/${TAG}_bench_$w.err; echo "bench $w rc=$?"
done
for w in reddit wikipedia scaled; do
  python bench.py --mode train --workload $w --steps 300 --warmup 20 --profile-steps 30 > $O/${TAG}_train_$w.json 2> $O/${TAG}_train_$w.err
  echo "train $w rc=$?"
done
python bench.py --impl reference --steps 20 --warmup 3 > $O/${TAG}_reference_reddit.json 2> $O/${TAG}_reference_reddit.err; echo "reference rc=$?"
# launch lists (warm L2: weights / memories stay resident between kernels as in the real step)
B="python bench.py --workload wikipedia --steps 20 --warmup 5 --cpu-batches 0 --profile-steps 0 --no-e2e --train-steps 0"
$B > $O/plain_a.log 2>&1 && ncu --cache-control none --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
  --log-file $O/${TAG}_launches_wikipedia_infer.csv $B > $O/ncu_a.log 2>&1
B="python bench.py --workload reddit --steps 20 --warmup 5 --cpu-batches 0 --profile-steps 0 --no-e2e --train-steps 0"
$B > $O/plain_b.log 2>&1 && ncu --cache-control none --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv \
  --log-file $O/${TAG}_launches_reddit_infer.csv $B > $O/ncu_b.log 2>&1
export TIGER_TRAIN_EAGER=1     # eager launches of the training step under ncu (same kernels as the graph replay)
for w in wikipedia reddit; do
  SKIP=3600; if [ $w = reddit ]; then SKIP=1200; fi     # ~30 warm-up steps of the workload's launches
  B="python bench.py --mode train --workload $w --steps 8 --warmup 30 --profile-steps 0 --cpu-batches 0 --no-e2e"
  $B > $O/plain_c.log 2>&1 && ncu --cache-control none --metrics gpu__time_duration.sum --clock-control none --launch-skip $SKIP -c 1000 --csv \
    --log-file $O/${TAG}_launches_${w}_train.csv $B > $O/ncu_c.log 2>&1
done
unset TIGER_TRAIN_EAGER
if [ "$MICRO" = 1 ]; then python bench.py --micro > $O/${TAG}_micro.json 2> $O/${TAG}_micro.err; echo micro_rc=$?; fi
if [ "$FULL" != 0 ]; then bash tools/gpu_ncu_full_train.sh; fi
ls -la $O | tail -30
