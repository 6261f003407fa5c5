"""Top stalled SASS instructions of one kernel in an .ncu-rep (needs --import-source / -lineinfo)."""
import collections, csv, io, subprocess, sys
rep, kernel = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--kernel-name', kernel], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[1]
ia, isrc, isamp, iex = h.index('Address'), h.index('Source'), h.index('# Samples'), h.index('Instructions Executed')
def I(x):
    try: return int(x)
    except ValueError: return 0
sec, n = [], 0
for r in rows:
    if r and r[0] == 'Kernel Name':
        n += 1
        continue
    if n == 1 and len(r) > isamp and r[ia] != 'Address':
        sec.append(r)
stalls = [c for c in h if c.startswith('stall_') and 'Not Issued' not in c]
agg = collections.Counter()
for r in sec:
    for s in stalls:
        agg[s] += I(r[h.index(s)])
tot = sum(I(r[isamp]) for r in sec)
print('samples', tot, [(k, v) for k, v in agg.most_common(6)])
for r in sorted(sec, key=lambda r: -I(r[isamp]))[:top]:
    st = sorted(((s, I(r[h.index(s)])) for s in stalls), key=lambda x: -x[1])[:2]
    print(f'{I(r[isamp]):5d} {I(r[iex]):8d}  {r[isrc][:90]:90s} {[x for x in st if x[1] > 0]}')
