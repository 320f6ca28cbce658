#!/usr/bin/env python
"""Development aid: per-kernel launch count / mean duration / share from an `ncu --metrics gpu__time_duration.sum --csv` log."""
import csv, sys, collections
rows = list(csv.reader(l for l in open(sys.argv[1], errors="replace") if l.startswith('"')))
h = rows[0]
ki, mi, vi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= vi or r[mi] != "gpu__time_duration.sum":
        continue
    a = agg.setdefault(r[ki], [0, 0.0])
    a[0] += 1; a[1] += float(r[vi].replace(",", ""))
tot = sum(a[1] for a in agg.values()) or 1.0
for k, (n, t) in agg.items():
    print(f"{k[:70]:70s} launches={n:4d} avg={t / n / 1e3:9.1f} us  share={100 * t / tot:5.1f}%")
