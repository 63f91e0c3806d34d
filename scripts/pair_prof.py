#!/usr/bin/env python
"""Summarise an ncu source-page CSV (--page source --csv --print-source sass) of one kernel:
per-instruction execution counts normalised to the hottest instruction, stall samples, and the
headline raw metrics.  usage: pair_prof.py <src.csv> <raw.csv> [min_samples]"""
import csv, sys
src, raw = sys.argv[1], sys.argv[2]
mins = int(sys.argv[3]) if len(sys.argv) > 3 else 250
rows = list(csv.reader(open(raw)))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'launch__registers_per_thread', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'smsp__warps_eligible.avg.per_cycle_active', 'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sectors_op_red.sum']
for r in rows[2:]:
    for i, h in enumerate(hdr):
        if h in keys or h.startswith('smsp__average_warps_issue_stalled') and float(r[i] or 0) > 0.15:
            print(f'{h} [{units[i]}] {r[i]}')
rows = list(csv.reader(open(src)))
hdr = rows[1]
si, ie, te, ns = hdr.index('Source'), hdr.index('Instructions Executed'), hdr.index('Thread Instructions Executed'), hdr.index('# Samples')
seen, out = set(), []
for r in rows[2:]:
    if len(r) <= ie or not r[ie].isdigit() or r[0] in seen:
        continue
    seen.add(r[0])
    out.append((r[0][-5:], r[si].strip(), int(r[ie]), int(r[te]), int(r[ns])))
mx = max(o[2] for o in out)
tot = sum(o[2] for o in out)
ts = sum(o[4] for o in out)
print(f'warp-instructions {tot}, static {len(out)}, hottest {mx}, samples {ts}')
with open('/tmp/t/pair_prof.txt', 'w') as f:
    for a, s, n, t, sa in out:
        f.write(f'{a} {n / mx:6.3f} {t / max(n, 1):5.1f} {sa:5d}  {s}\n')
for a, s, n, t, sa in out:
    if sa >= mins:
        print(f'{a} {n / mx:6.3f} {t / max(n, 1):5.1f} {sa:5d} {100 * sa / ts:4.1f}%  {s}')
