# quick check after a kernel change: GPU suites + the two seq benches
python -m pytest tests/test_train_gpu.py tests/test_ops_gpu.py tests/test_engine_gpu.py tests/test_dropin_gpu.py -m gpu -x -q 2>&1 | tail -5
for w in wikipedia mooc; do
  python bench.py --workload $w --cpu-batches 20 --parity-batches 1 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$w',d['ms_per_step'],d['e2e']['value'],d.get('parity_checked'), (d.get('train_step') or {}).get('ms_per_step'))"
done
python bench.py --workload wikipedia --mode train --steps 100 --warmup 10 --cpu-batches 10 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('train wikipedia',d['ms_per_step'],d['e2e']['value'],d.get('parity_checked'))"
