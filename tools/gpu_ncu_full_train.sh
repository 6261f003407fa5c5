#!/bin/bash
# ncu --set full of the training step's heavy kernels (wikipedia-shaped, seq restarter), after the same command ran clean
O=gpurun_out
B="python bench.py --mode train --workload wikipedia --steps 40 --warmup 30 --profile-steps 0 --cpu-batches 0 --no-e2e"
$B > $O/r02_full_plain.json 2> $O/r02_full_plain.err || { tail -5 $O/r02_full_plain.err; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"train_seq_pool|gemm_tf32x3_kernel|train_attn_core|train_mse|train_score_head" \
  --launch-skip 2400 -c 60 -f -o $O/prof_r02_train_wikipedia $B > $O/r02_full_ncu.log 2>&1
tail -2 $O/r02_full_ncu.log
ls -la $O/prof_r02_train_wikipedia.ncu-rep
python tools/ncu_summary.py $O/prof_r02_train_wikipedia.ncu-rep > $O/prof_r02_train_wikipedia.md 2>&1
grep -c "^## " $O/prof_r02_train_wikipedia.md
