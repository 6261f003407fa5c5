#!/bin/bash
# GPU check of the persistent / general GEMM + the seq-restarter bench (used during round 2)
python -m pytest tests/test_ops_gpu.py -x -q -k "sgemm or seq or restart" 2>&1 | tail -15
python -m pytest tests/test_engine_gpu.py -x -q 2>&1 | tail -5
python bench.py --workload wikipedia --steps 500 --warmup 20 --cpu-batches 0 > gpurun_out/r02b_wiki.json 2> gpurun_out/r02b_wiki.err
tail -c 300 gpurun_out/r02b_wiki.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02b_wiki.json").read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"])
for k, v in d["kernels"].items():
    print(k, round(v["us"], 1), v["launches_per_step"])
PY
