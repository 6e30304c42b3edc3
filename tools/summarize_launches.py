#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (share of total time)."""
import csv, sys, re, collections
rows = [r for r in csv.reader(l for l in open(sys.argv[1]) if l.startswith('"'))]
hdr = rows[0]; ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
agg = collections.OrderedDict(); tot = 0.0
for r in rows[1 + skip:]:
    name = re.sub(r"\(.*", "", r[ki]); name = re.sub(r"^void ", "", name)
    v = float(r[vi].replace(",", "")); v = v / 1000.0 if r[ui] == "ns" else v   # -> us
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v; tot += v
print(f"{'kernel':70s} {'launches':>8s} {'total us':>10s} {'share':>7s}")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} {n:8d} {v:10.1f} {100 * v / tot:6.1f}%")
print(f"{'TOTAL':70s} {sum(a[0] for a in agg.values()):8d} {tot:10.1f}")
