#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source cuda,sass` dump: executed warp
instructions and stall samples per CUDA source line (needs -lineinfo builds)."""
import csv
import sys

path, n_units = sys.argv[1], float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = list(csv.reader(open(path)))
hdr = next(r for r in rows if "Instructions Executed" in r)
ix, sx = hdr.index("Instructions Executed"), hdr.index("# Samples")
items, total, tot_s = [], 0, 0
for r in rows:
    if len(r) <= ix or not r[0].isdigit():
        continue
    try:
        n, s = int(r[ix]), int(r[sx])
    except ValueError:
        continue
    total += n
    tot_s += s
    items.append((n, s, r[0], r[1].strip()[:100]))
print(f"total warp instructions {total}  ({total / n_units:.1f} per unit), samples {tot_s}")
for n, s, line, src in sorted(items, reverse=True)[:int(sys.argv[3]) if len(sys.argv) > 3 else 40]:
    print(f"{n / total * 100:5.1f}% inst {s / max(tot_s, 1) * 100:5.1f}% smp {n / n_units:8.1f}/unit  L{line}: {src}")
