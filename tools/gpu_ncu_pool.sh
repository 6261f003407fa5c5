#!/bin/bash
# ncu --set full of the seq pooling kernel inside the inference step (wikipedia shape, ~23 restarted nodes): details page +
# the source lines with the most stall samples.  The command runs clean without ncu first.
O=gpurun_out
B="python bench.py --workload wikipedia --steps 20 --warmup 5 --cpu-batches 0 --profile-steps 0 --no-e2e --train-steps 0"
$B > $O/pool_plain.log 2>&1 || { tail -3 $O/pool_plain.log; exit 1; }
ncu --set full --import-source on --clock-control none -k regex:train_seq_pool_kernel --launch-skip 20 -c 1 -f -o /tmp/pool $B > $O/pool_ncu.log 2>&1
tail -2 $O/pool_ncu.log
ncu -i /tmp/pool.ncu-rep --page details 2>/dev/null | grep -E "Duration|Elapsed Cycles|Executed Ipc|Issue Slots Busy|No Eligible|Eligible Warps|Stall|L1/TEX Hit|L2 Hit|Registers|Theoretical Occ|Achieved Occ|Shared Memory Config|Bank|Mem Busy|Max Bandwidth|Warp Cycles Per Issued|Est. Speedup|uncoalesced|excessive" | head -50 > $O/pool_details.txt
ncu -i /tmp/pool.ncu-rep --page source --csv 2>/dev/null > /tmp/pool_src.csv
python - <<'PY'
import csv
rows = list(csv.reader(open('/tmp/pool_src.csv', errors='replace')))
h = next(i for i, r in enumerate(rows) if 'Source' in r or '# Samples' in ' '.join(r) or 'Sampling' in ' '.join(r))
hdr = rows[h]
print(hdr[:12])
si = next((i for i, c in enumerate(hdr) if c.startswith('# Samples') or c == 'Warp Stall Sampling (All Samples)' or 'Samples' in c), None)
src = hdr.index('Source') if 'Source' in hdr else 1
data = []
for r in rows[h + 1:]:
    try:
        data.append((float(r[si].replace(',', '') or 0), r[src][:150]))
    except Exception:
        pass
tot = sum(d[0] for d in data) or 1
out = open('gpurun_out/pool_source_top.txt', 'w')
for s, l in sorted(data, key=lambda x: -x[0])[:30]:
    out.write(f'{100 * s / tot:5.1f}%  {l}\n')
out.close()
PY
cat $O/pool_details.txt | head -40; cat $O/pool_source_top.txt
