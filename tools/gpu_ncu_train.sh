#!/bin/bash
# ncu launch list (gpu__time_duration) of a short train-mode bench; $1 = workload
W=${1:-reddit}
python bench.py --mode train --workload $W --steps 6 --warmup 3 --profile-steps 0 --cpu-batches 0 --no-e2e > gpurun_out/r02_ncu_train_$W.plain.json 2> gpurun_out/r02_ncu_train_$W.plain.err || { tail -5 gpurun_out/r02_ncu_train_$W.plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 600 -c 700 --csv --log-file gpurun_out/r02_ncu_train_$W.csv \
  python bench.py --mode train --workload $W --steps 6 --warmup 3 --profile-steps 0 --cpu-batches 0 --no-e2e > gpurun_out/r02_ncu_train_$W.log 2>&1
tail -3 gpurun_out/r02_ncu_train_$W.log
python - <<PY
import csv, collections
rows=[r for r in csv.reader(open('gpurun_out/r02_ncu_train_$W.csv')) if len(r)>10]
hdr=rows[0]; ik=hdr.index('Kernel Name'); iv=hdr.index('Metric Value')
agg=collections.OrderedDict()
for r in rows[1:]:
    k=r[ik][:60]
    a=agg.setdefault(k,[0,0.0]); a[0]+=1; a[1]+=float(r[iv].replace(',',''))/1e3
tot=sum(a[1] for a in agg.values())
print('total us', tot, 'launches', sum(a[0] for a in agg.values()))
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][1])[:25]: print('%-62s n=%4d  %9.1f us  avg %7.1f'%(k,a[0],a[1],a[1]/a[0]))
PY
