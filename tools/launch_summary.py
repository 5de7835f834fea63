"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): python tools/launch_summary.py FILE.csv [passes]"""
import collections
import csv
import sys

rows = list(csv.reader(l for l in open(sys.argv[1]) if not l.startswith("==")))
passes = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    a = agg.setdefault(r[ki][:100], [0, 0.0, 0.0])
    a[0] += 1; a[1] += v; a[2] = max(a[2], v)
tot = sum(a[1] for a in agg.values())
print(f"total {tot / 1e6 / passes:.3f} ms per pass ({passes:g} passes in the list)")
for k, (n, t, mx) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"{t / 1e6 / passes:9.3f} ms/pass  {n:4d} launches  max {mx / 1e6:8.3f} ms  {k}")
