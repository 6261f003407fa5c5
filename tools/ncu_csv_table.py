"""Aggregates an ncu CSV log (`ncu --metrics ... --csv --log-file X`, one line per launch and metric) into one markdown
row per (kernel, grid): launches, mean duration, mean DRAM bytes read / written per launch and the pipe percentages.
usage: python tools/ncu_csv_table.py X.csv [title]"""
import collections
import csv
import sys

SCALE = {'ns': 1e-3, 'us': 1.0, 'ms': 1e3, 'byte': 1.0, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}


def table(path):
    rows = list(csv.reader(open(path, errors='replace')))
    hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
    h = rows[hi]
    agg = collections.OrderedDict()
    for r in rows[hi + 1:]:
        if len(r) < len(h):
            continue
        d = dict(zip(h, r))
        key = (d['Kernel Name'].split('(')[0].replace('void ', ''), d['Grid Size'], d['Block Size'])
        val = float(d['Metric Value'].replace(',', '')) * SCALE.get(d['Metric Unit'], 1.0)
        agg.setdefault(key, collections.defaultdict(list))[d['Metric Name']].append(val)
    out = ['| kernel | grid | block | launches | us/launch | DRAM read MB | DRAM write MB | DRAM GB/s | tensor pipe % | SM % | DRAM % |',
           '|---|---|---|---|---|---|---|---|---|---|---|']
    mean = lambda v, k: sum(v[k]) / len(v[k]) if v.get(k) else float('nan')
    for (name, grid, block), v in agg.items():
        us = mean(v, 'gpu__time_duration.sum')
        rd, wr = mean(v, 'dram__bytes_read.sum'), mean(v, 'dram__bytes_write.sum')
        out.append(f'| {name} | {grid} | {block} | {len(v["gpu__time_duration.sum"])} | {us:.1f} | {rd / 1e6:.2f} | {wr / 1e6:.2f} | '
                   f'{(rd + wr) / us / 1e3:.0f} | {mean(v, "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"):.1f} | '
                   f'{mean(v, "sm__throughput.avg.pct_of_peak_sustained_elapsed"):.1f} | '
                   f'{mean(v, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"):.1f} |')
    return '\n'.join(out)


if __name__ == '__main__':
    if len(sys.argv) > 2:
        print(f'## {sys.argv[2]}\n')
    print(table(sys.argv[1]))
