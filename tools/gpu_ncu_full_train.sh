#!/bin/bash
# ncu --set full of the training step's heavy kernels (wikipedia-shaped, seq restarter), after the same command ran clean.
# Only the text summaries travel back (gpurun_out is capped at 64 MiB; a --set full report of 60 launches is ~100 MB).
O=gpurun_out
export TIGER_TRAIN_EAGER=1
B="python bench.py --mode train --workload wikipedia --steps 40 --warmup 30 --profile-steps 0 --cpu-batches 0 --no-e2e"
$B > $O/r02_full_plain.json 2> $O/r02_full_plain.err || { tail -5 $O/r02_full_plain.err; exit 1; }
R=/tmp/prof_r02_train_wikipedia
ncu --set full --clock-control none -k regex:"train_seq_pool|gemm_tf32x3_kernel|gemm_pp_kernel|gemm_pp_pack|seq_tail_layer|train_attn_core|train_seq_tokens_bwd|train_seq_vbias_bwd|train_adam" \
  --launch-skip 2600 -c 56 -f -o $R $B > $O/r02_full_ncu.log 2>&1
tail -2 $O/r02_full_ncu.log
python tools/ncu_summary.py $R.ncu-rep > $O/prof_r02_train_wikipedia.md 2>&1
ncu -i $R.ncu-rep --page details --csv 2>/dev/null | grep -i "stall\|Issued Warp\|Eligible\|No Eligible\|Registers Per\|Achieved Occupancy\|Theoretical Occ\|Local\|Shared Memory Config\|Block Limit" | cut -d, -f5,13-16 | sort | uniq -c | sort -rn | head -80 > $O/prof_r02_train_wikipedia_details.txt
grep -c "^## " $O/prof_r02_train_wikipedia.md
