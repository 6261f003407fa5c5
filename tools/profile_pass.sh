#!/bin/bash
# Round-end measurement pass (run on the GPU box through gpurun): smoke, default bench, warm / cold ncu launch
# lists, one `ncu --set full` capture of the dense kernels, micro mode and its per-kernel captures.
# Every ncu command runs only after the identical command has exited 0 without ncu.
TAG=${1:-r01}
O=gpurun_out
python __graft_entry__.py smoke > $O/smoke_$TAG.log 2>&1; echo smoke_rc=$?
python bench.py > $O/bench_${TAG}_final.json 2> $O/bench_${TAG}_final.err; echo bench_rc=$?
B="python bench.py --steps 20 --warmup 5 --cpu-batches 0 --profile-steps 0 --no-e2e"
$B > $O/plain_b.log 2>&1 && ncu --cache-control none --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv \
  --log-file $O/launches_${TAG}_warm.csv $B > $O/ncu_l.log 2>&1
$B > $O/plain_b1.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv \
  --log-file $O/launches_${TAG}_cold.csv $B > $O/ncu_l2.log 2>&1
$B > $O/plain_b2.log 2>&1 && ncu --set full --clock-control none --import-source on \
  -k regex:"gru_update_kernel|gemm_tf32x3|attn_score_pool|link_score|compact_involved" -s 60 -c 7 -f -o $O/prof_${TAG}_dense $B > $O/ncu_f.log 2>&1
M="python bench.py --micro"
$M > $O/micro_${TAG}_final.json 2> $O/plain_m.err; echo micro_rc=$?
for k in gather_rows_kernel scatter_rows_kernel store_messages_kernel right_writeback_kernel left_writeback_kernel find_recent_kernel; do
  $M > /dev/null 2>&1 && ncu --set full --clock-control none -k regex:$k -s 10 -c 1 -f -o $O/prof_${TAG}_micro_$k $M > $O/ncu_m_$k.log 2>&1
done
echo done
