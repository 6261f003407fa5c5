"""Turns the files tools/profile_pass.sh left in gpurun_out/ into the tracked summaries under profiles/."""
import json, subprocess, sys, shutil, csv, io
tag = sys.argv[1]
step_ms = json.load(open(f'gpurun_out/bench_{tag}_final.json'))['ms_per_step']
run = lambda *a: subprocess.run([sys.executable, *a], capture_output=True, text=True).stdout
warm = run('tools/launch_table.py', f'gpurun_out/launches_{tag}_warm.csv')
cold = run('tools/launch_table.py', f'gpurun_out/launches_{tag}_cold.csv')
B = 'python bench.py --steps 20 --warmup 5 --cpu-batches 0 --profile-steps 0 --no-e2e'
open('profiles/r01_launches_reddit.md', 'w').write(f'''# Round 1 - ncu launch lists, reddit-shaped workload, batch 200 (final build of the round)

Command (`tools/profile_pass.sh`, after the same command exited 0 without ncu): `ncu [--cache-control none] --metrics gpu__time_duration.sum --clock-control none -c 1400 --csv --log-file gpurun_out/launches_{tag}_{{warm,cold}}.csv {B}`; tables made by `tools/launch_table.py` (mean over the last graph-replayed steps).

Per-launch times under ncu are serialised and exclude every overlap the real step has: the finder runs on the copy-in stream beside the previous batch, the restarter beside the GRU, select / right write-back / message store beside the attention chain, the left write-back beside the link scorer, and programmatic dependent launch overlaps each kernel's set-up (and the score_pool gathers) with its predecessor.  The SHARE column is what is comparable with bench.py, whose graph-replayed step takes {step_ms:.4f} ms - less than the serialised total below.

`gemm_tf32x3_ts_kernel` with grid (85,1,1): the Wqk product F1 (center rows gathered by the producers) and the fc2 + link-scorer-fold product F3; grid (30,4,1): the W2f product F2, K split over a 4-CTA cluster.

## warm L2 (--cache-control none: weights, memories and the message store stay L2 resident between kernels, as in the real step)

{warm.split(chr(10), 1)[1]}
## cold caches (ncu default: caches flushed before every kernel)

{cold.split(chr(10), 1)[1]}''')
full = run('tools/ncu_summary.py', f'gpurun_out/prof_{tag}_dense.ncu-rep')
open('profiles/r01_ncu_full_dense_kernels.md', 'w').write(f'''# Round 1 - `ncu --set full` of the dominant kernels (reddit-shaped workload, B=200, final build of the round)

Command (`tools/profile_pass.sh`): `ncu --set full --clock-control none --import-source on -k regex:"gru_update_kernel|gemm_tf32x3|attn_score_pool|link_score|compact_involved" -s 60 -c 7 -o gpurun_out/prof_{tag}_dense {B}` (same command exited 0 without ncu first). Values read with `ncu -i ... --page raw --csv` (`tools/ncu_summary.py`); per-instruction stall reasons with `tools/ncu_stalls.py`. Caches are flushed before each profiled kernel, so durations are cold-cache.

The `gemm_tf32x3_ts_kernel` captures are told apart by their grid: 120 CTAs = F2 (W2f, 4-CTA cluster split-K), 85 CTAs with ~1.3 MB read = F3 (fc2 + scorer fold), 85 CTAs with ~2.1 MB read = F1 (Wqk, gathered center rows).
''' + full)
# traffic.json from the dense capture
raw = subprocess.run(['ncu', '-i', f'gpurun_out/prof_{tag}_dense.ncu-rep', '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
h, units = rows[0], rows[1]
def col(r, m):
    v = float(r[h.index(m)].replace(',', ''))
    u = units[h.index(m)]
    return v * {'byte': 1, 'Kbyte': 1e3, 'Mbyte': 1e6, 'Gbyte': 1e9}[u]
att, gru, parts = 0.0, 0.0, []
for r in rows[2:]:
    name = r[h.index('Kernel Name')]
    b = col(r, 'dram__bytes_read.sum') + col(r, 'dram__bytes_write.sum')
    if 'gru_update' in name: gru = b
    if 'gemm_tf32x3_ts' in name or 'attn_score_pool' in name:
        att += b
        parts.append(int(b))
json.dump({'_doc': 'dram__bytes_read.sum + dram__bytes_write.sum per launch from the ncu --set full capture summarised in profiles/r01_ncu_full_dense_kernels.md (bench.py copies the entry of the dominant entry point into roofline.traffic); tiger_temporal_attention = sum over its 4 kernels ' + ' + '.join(map(str, parts)),
           'reddit': {'tiger_temporal_attention': int(att), 'tiger_gru_update': int(gru)}}, open('profiles/traffic.json', 'w'), indent=1)
# micro
micro = json.load(open(f'gpurun_out/micro_{tag}_final.json'))
out = [f'''# Round 1 - `ncu --set full` of the gather / scatter / search kernels in micro mode (262,144 rows, tables >> L2)

Commands (`tools/profile_pass.sh`): `python bench.py --micro` exited 0, then `ncu --set full --clock-control none -k regex:<kernel> -s 10 -c 1 -o gpurun_out/prof_{tag}_micro_<kernel> python bench.py --micro` per kernel. `gpu__dram_throughput` is relative to the hardware peak ncu assumes (~8.2 TB/s); `bench.py --micro` reports against the MEASURED copy peak of {micro["peak"]:.0f} GB/s (MEASURED_PEAKS.json).
''']
for k in ['gather_rows_kernel', 'scatter_rows_kernel', 'store_messages_kernel', 'right_writeback_kernel', 'left_writeback_kernel', 'find_recent_kernel']:
    out.append(run('tools/ncu_summary.py', f'gpurun_out/prof_{tag}_micro_{k}.ncu-rep'))
out.append('\n## bench.py --micro (CUDA events, 20 launches after 3 warm-ups; same build)\n\n```\n')
for k, v in micro['micro'].items():
    out.append(f'{k:26s} {v["us"]:8.1f} us {v["gbs"]:8.0f} GB/s  {v["frac_of_peak"]:.3f} of measured peak   {v["note"]}\n')
out.append('```\n')
open('profiles/r01_ncu_full_micro_gather_scatter.md', 'w').write(''.join(out))
shutil.copy(f'gpurun_out/bench_{tag}_final.json', 'profiles/r01_bench_reddit_final.json')
shutil.copy(f'gpurun_out/micro_{tag}_final.json', 'profiles/r01_micro_final.json')
print('ok')
