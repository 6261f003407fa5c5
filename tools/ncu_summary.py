"""Prints the roofline-relevant metrics of every kernel in an .ncu-rep (ncu -i ... --page raw --csv)."""
import csv, io, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h = rows[0]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'lts__t_sector_hit_rate.pct', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'launch__waves_per_multiprocessor']
units = rows[1]
for r in rows[2:]:
    name = r[h.index('Kernel Name')].split('(')[0]
    print(f'\n## {name}\n\n| metric | value | unit |\n|---|---|---|')
    for m in want:
        if m in h:
            print(f'| {m} | {r[h.index(m)]} | {units[h.index(m)]} |')
