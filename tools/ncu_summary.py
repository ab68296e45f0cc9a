"""Condensed view of an ncu report exported with --page raw --csv and --page source --csv.
usage: ncu_summary.py RAW.csv [SRC.csv]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
vals = rows[2] if len(rows) > 2 else rows[1]
want = ['gpu__time_duration.sum', 'sm__throughput.avg.pct', 'smsp__issue_active.avg.pct', 'sm__warps_active.avg.pct',
        'launch__registers_per_thread', 'launch__occupancy_limit', 'smsp__inst_executed.sum', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'launch__grid_size', 'launch__block_size', 'sm__inst_executed_pipe_tensor',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'dram__throughput.avg.pct']
for h, v in zip(hdr, vals):
    if any(w in h for w in want) and 'per_second' not in h:
        print(f"{h:80s} {v}")
stalls = [(float(v), h) for h, v in zip(hdr, vals) if 'issue_stalled' in h and 'per_issue_active' in h]
for v, h in sorted(stalls, reverse=True)[:8]:
    print(f"  stall {h.split('issue_stalled_')[1].split('_per_issue')[0]:24s} {v:.2f}")
if len(sys.argv) > 2:
    rows = list(csv.reader(open(sys.argv[2])))
    hdr = rows[1]
    ci = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    groups = []
    for i, r in enumerate(data):
        n = int(r[ci['Instructions Executed']] or 0)
        if groups and groups[-1][2] == n:
            groups[-1][1] = i
        else:
            groups.append([i, i, n])
    tot = sum((g[1] - g[0] + 1) * g[2] for g in groups)
    print('sass instructions', len(data), 'executed', tot)
    for g in groups:
        c = (g[1] - g[0] + 1) * g[2]
        if c > tot * 0.02:
            ops = {}
            samples = 0
            for r in data[g[0]:g[1] + 1]:
                src = r[ci['Source']].split()
                op = src[1] if src[0].startswith('@') else src[0]
                op = op.split('.')[0]
                ops[op] = ops.get(op, 0) + 1
                samples += int(r[ci['# Samples']] or 0)
            top = sorted(ops.items(), key=lambda x: -x[1])[:7]
            print(f"  [{g[0]:4d}-{g[1]:4d}] x{g[2]:<8d} {100 * c / tot:5.1f}% samples {samples:5d}  {top}")
