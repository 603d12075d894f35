"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): launches, mean duration and share per kernel.
usage: python tools/launch_summary.py launches.csv ["header line"]"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if r]
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
h = rows[hdr]
kn, mv, mu = h.index("Kernel Name"), h.index("Metric Value"), h.index("Metric Unit")
acc = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= mv:
        continue
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    u = r[mu]
    us = v / 1e3 if u in ("ns", "nsecond") else (v if u in ("us", "usecond") else v * 1e3 if u in ("ms", "msecond") else v)
    a = acc.setdefault(r[kn], [0, 0.0])
    a[0] += 1
    a[1] += us
tot = sum(a[1] for a in acc.values())
if len(sys.argv) > 2:
    print(sys.argv[2])
for k, (n, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:60]:60s} n={n:3d} avg={t / n:10.1f} us share={t / tot:.4f}")
