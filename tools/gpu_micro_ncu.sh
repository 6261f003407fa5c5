#!/bin/bash
# bench.py --micro (HBM / tensor fractions of the gather / scatter / search / GRU / attention kernels at >= 64k rows over
# 1M-row tables), then the same command under ncu for the DRAM bytes per launch (after it exited 0 without ncu).
TAG=${1:-r02}
O=gpurun_out
B="python bench.py --micro"
$B > $O/${TAG}_micro.json 2> $O/${TAG}_micro.err || { tail -5 $O/${TAG}_micro.err; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
  --clock-control none -k regex:"gather_rows|scatter_rows|store_messages|writeback|find_recent|gru_update|attn_|gemm_tf32x3_ts" \
  --csv --log-file $O/${TAG}_micro_ncu.csv $B > $O/${TAG}_micro_ncu.log 2>&1
echo ncu_rc=$?; wc -l $O/${TAG}_micro_ncu.csv
python - <<'PY'
import json
d=json.loads(open('gpurun_out/%s_micro.json' % __import__('sys').argv[1] if len(__import__('sys').argv)>1 else 'gpurun_out/r02_micro.json').read().strip().splitlines()[-1])
for k,v in d['micro'].items(): print(k, v['rows'], round(v['us'],1), round(v['gbs']), round(v['frac_of_peak'],3), v.get('tensor',{}).get('frac'))
PY
