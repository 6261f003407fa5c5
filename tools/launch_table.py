"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list into a per-kernel table
(mean over the last graph-replayed steps)."""
import collections, csv, sys
path = sys.argv[1]
tail = float(sys.argv[2]) if len(sys.argv) > 2 else 0.33
rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
h = rows[0]
ki, vi, gi, bi = h.index('Kernel Name'), h.index('Metric Value'), h.index('Grid Size'), h.index('Block Size')
data = rows[1:]
seq = data[-int(len(data) * tail):]
agg = collections.OrderedDict()
for r in seq:
    k = (r[ki].split('(')[0], r[gi], r[bi])
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(',', ''))
steps = min(a[0] for a in agg.values())
tot = sum(a[1] for a in agg.values())
print(f'{len(data)} launches, table over the last {len(seq)} ({steps} steps)')
print('| kernel | launches/step | grid | block | us/launch | us/step | share |')
print('|---|---|---|---|---|---|---|')
for k, a in agg.items():
    print(f'| {k[0]} | {a[0] / steps:.1f} | {k[1]} | {k[2]} | {a[1] / a[0] / 1e3:.2f} | {a[1] / steps / 1e3:.1f} | {100 * a[1] / tot:.1f}% |')
print(f'| **total** | | | | | {tot / steps / 1e3:.1f} | |')
