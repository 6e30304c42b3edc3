#!/usr/bin/env python
"""Summarise `ncu -i X.ncu-rep --page source --csv` (SASS view): stall-sample totals by reason and the hottest instructions.

    ncu -i prof.ncu-rep --page source --csv > src.csv ; python tools/ncu_hotspots.py src.csv [top_n]
"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
col = {n: i for i, n in enumerate(hdr)}
stall_cols = [n for n in hdr if n.startswith("stall_") and "Not Issued" not in n]
data = []
for r in rows[2:]:
    if len(r) < len(hdr):
        continue
    try:
        samples = int(r[col["# Samples"]] or 0)
    except ValueError:
        continue
    data.append((samples, r))
total = sum(s for s, _ in data)
print(f"total samples {total}, instructions {len(data)}")
agg = {n: sum(int(r[col[n]] or 0) for _, r in data) for n in stall_cols}
for n, v in sorted(agg.items(), key=lambda kv: -kv[1]):
    if v:
        print(f"  {n:28s} {v:8d} {100.0 * v / max(total, 1):5.1f} %")
print("hottest instructions:")
for idx, (s, r) in enumerate(data):
    r.append(idx)
for s, r in sorted(data, key=lambda t: -t[0])[:top]:
    reasons = sorted(((int(r[col[n]] or 0), n) for n in stall_cols), reverse=True)[:2]
    print(f"  #{r[-1]:5d} {s:6d} {100.0 * s / max(total, 1):5.1f} %  {r[col['Source']][:70]:70s} {reasons[0][1]}={reasons[0][0]} {reasons[1][1]}={reasons[1][0]}")
