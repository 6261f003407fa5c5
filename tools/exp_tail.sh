# experiment: small-R tail of the seq restarter on / off (inference + training), same box
python -m pytest tests/test_train_gpu.py tests/test_ops_gpu.py tests/test_engine_gpu.py tests/test_dropin_gpu.py -m gpu -x -q 2>&1 | tail -5
for w in mooc wikipedia; do
for v in "TIGER_SEQ_TAIL_ROWS=0" "TIGER_SEQ_TAIL_ROWS=64" "TIGER_SEQ_TAIL_ROWS=64 TIGER_SEQ_TAIL_SERIAL=1"; do
  env $v python bench.py --workload $w --steps 300 --warmup 30 --cpu-batches 20 --parity-batches 1 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('$w','$v',d['ms_per_step'],d['e2e']['value'],d.get('parity_checked'))"
done; done
for v in "TIGER_SEQ_TAIL_ROWS=0" "TIGER_SEQ_TAIL_ROWS=64"; do
  env $v python bench.py --workload wikipedia --mode train --steps 100 --warmup 10 --cpu-batches 10 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('train wikipedia','$v',d['ms_per_step'],d['e2e']['value'],d.get('parity_checked'))"
done
