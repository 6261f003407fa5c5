"""Per-step table of an `ncu --metrics gpu__time_duration.sum --csv` launch list: steps are delimited by the launches of a
marker kernel (the 3B-query neighbor finder that opens every batch), the table is the mean over the complete steps found
at the END of the capture (steady state)."""
import collections, csv, sys
path, marker = sys.argv[1], sys.argv[2]
per_step = int(sys.argv[3]) if len(sys.argv) > 3 else 1          # marker launches per step
keep = int(sys.argv[4]) if len(sys.argv) > 4 else 4              # complete steps averaged
rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
h = rows[0]
ki, vi, gi, bi = h.index('Kernel Name'), h.index('Metric Value'), h.index('Grid Size'), h.index('Block Size')
data = rows[1:]
marks = [i for i, r in enumerate(data) if r[ki].startswith(marker)][::per_step]
marks = marks[-(keep + 1):]
seq = data[marks[0]:marks[-1]]
steps = len(marks) - 1
agg = collections.OrderedDict()
for r in seq:
    k = (r[ki].split('(')[0][:60], r[gi], r[bi])
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += float(r[vi].replace(',', ''))
tot = sum(a[1] for a in agg.values())
print(f'{len(data)} launches captured; table = mean over the last {steps} complete steps ({len(seq)} launches)\n')
print('| kernel | launches/step | grid | block | us/launch | us/step | share |')
print('|---|---|---|---|---|---|---|')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'| {k[0]} | {a[0] / steps:.1f} | {k[1]} | {k[2]} | {a[1] / a[0] / 1e3:.2f} | {a[1] / steps / 1e3:.1f} | {100 * a[1] / tot:.1f}% |')
print(f'| **total** | {len(seq) / steps:.1f} | | | | {tot / steps / 1e3:.1f} | |')
