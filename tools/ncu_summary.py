#!/usr/bin/env python3
"""Summarise an ncu report exported as CSV (--page raw / --page source): key metrics,
instructions / shared wavefronts / stall samples per barrier-delimited code region."""
import csv, sys, collections
raw, src, rows_per = sys.argv[1], sys.argv[2], float(sys.argv[3]) if len(sys.argv) > 3 else 3.2e7
rows = list(csv.reader(open(raw)))
m = {h: (u, v) for h, u, v in zip(rows[0], rows[1], rows[2])}
keys = ['gpu__time_duration.sum', 'sm__cycles_elapsed.avg.per_second', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'sm__inst_executed.sum.per_cycle_active', 'sm__issue_active.avg.pct_of_peak_sustained_elapsed',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sass__inst_executed_local_loads', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__block_size', 'launch__grid_size', 'launch__shared_mem_per_block_dynamic']
keys += sorted(k for k in m if 'issue_stalled' in k and 'per_issue_active' in k)
for k in keys:
    if k in m:
        print(f"{k:95s} {m[k][0]:12s} {m[k][1]}")
rows = list(csv.reader(open(src)))
hdr = rows[1]
isrc, iex, ismp, iw = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("# Samples"), hdr.index("L1 Wavefronts Shared")
data = [(int(r[iex]), int(r[ismp]), r[isrc].strip(), int(r[iw] or 0)) for r in rows[2:] if len(r) > iex]
tot = sum(d[0] for d in data); ts = sum(d[1] for d in data)
# the source page can count more than the kernel executed (it did by 1.42x on k_main_tc): scale the
# per-instruction counts so that they add up to smsp__inst_executed.sum of the raw page
ex_raw = float(m['smsp__inst_executed.sum'][1]) if 'smsp__inst_executed.sum' in m else tot
scale = ex_raw / tot if tot else 1.0
print(f"\nsmsp__inst_executed.sum {ex_raw:.4e}; source-page sum {tot:.4e} (counts below scaled by {scale:.3f})")
data = [(e * scale, s_, srcl, w) for e, s_, srcl, w in data]
tot = ex_raw
print(f"\nwarp instructions per obs-row: {tot / rows_per:.1f}   shared wavefronts per obs-row: {sum(d[3] for d in data) / rows_per:.1f}")
start = ex = sm = wf = 0
for i, (e, s_, srcl, w) in enumerate(data):
    ex += e; sm += s_; wf += w
    if 'BAR.' in srcl or srcl.startswith('EXIT'):
        if ex / rows_per > 0.5 or sm / ts > 0.01:
            print(f"  sass {start:5d}-{i:5d}  {ex / rows_per:8.1f} instr/obs-row {wf / rows_per:7.1f} wf/obs-row {sm / ts * 100:5.1f}% of stall samples   ends at {srcl[:40]}")
        start, ex, sm, wf = i + 1, 0, 0, 0
op = collections.Counter()
for e, s_, srcl, w in data:
    t = srcl.split()
    o = t[1] if t[0].startswith('@') else t[0]
    op[o.split('.')[0]] += e
print("  opcode mix per obs-row:", {k: round(v / rows_per, 1) for k, v in op.most_common(22)})
