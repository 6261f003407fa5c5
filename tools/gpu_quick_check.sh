#!/bin/bash
# Quick check on one B200 after a change (through tools/gpurun_retry.sh): smoke, the GPU suites, the default bench line.
O=gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -2
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > $O/quick_bench.json 2> $O/quick_bench.err; echo "bench rc=$?"
tail -1 $O/quick_bench.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); t=d.get('train_step') or {}
print('reddit', d['ms_per_step'], d['e2e']['value'], d.get('parity_checked'), 'train_step', t.get('ms_per_step'), t.get('cuda_graph'), t.get('cuda_graph_error'))"
