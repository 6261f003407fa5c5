O=gpurun_out
python -m pytest tests/test_ops_gpu.py tests/test_engine_gpu.py tests/test_train_gpu.py -m gpu -x -q -k "seq or wikipedia or full or baseline or graph" 2>&1 | tail -3
for w in wikipedia mooc; do
  python bench.py --workload $w > $O/r02_bench_$w.json 2> $O/r02_bench_$w.err; echo "bench $w rc=$?"
  tail -1 $O/r02_bench_$w.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$w',d['ms_per_step'],d['e2e']['value'],d.get('parity_checked'), (d.get('train_step') or {}).get('ms_per_step'))"
done
