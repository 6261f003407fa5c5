# after the vectorised tail kernel: suites, then the seq-restarter inference lines + launch list again
O=gpurun_out
python -m pytest tests/test_train_gpu.py tests/test_ops_gpu.py tests/test_engine_gpu.py tests/test_dropin_gpu.py -m gpu -x -q 2>&1 | tail -4
for w in wikipedia mooc lastfm scaled; do
  python bench.py --workload $w > $O/r02_bench_$w.json 2> $O/r02_bench_$w.err; echo "bench $w rc=$?"
  tail -1 $O/r02_bench_$w.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$w',d['ms_per_step'],d['e2e']['value'],d.get('parity_checked'), (d.get('train_step') or {}).get('ms_per_step'))"
done
B="python bench.py --workload wikipedia --steps 20 --warmup 5 --cpu-batches 0 --profile-steps 0 --no-e2e --train-steps 0"
$B > $O/plain_a.log 2>&1 && ncu --cache-control none --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv \
  --log-file $O/r02_launches_wikipedia_infer.csv $B > $O/ncu_a.log 2>&1
echo ncu_rc=$?
