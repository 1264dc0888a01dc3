"""Summarise an `ncu --metrics gpu__time_duration.sum,launch__grid_size --csv` launch list per (kernel, grid size).
usage: launch_summary_grid.py <csv>"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ii, ki, mi, vi = hdr.index("ID"), hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
per = collections.OrderedDict()
for r in rows[1:]:
    e = per.setdefault(r[ii], {"k": r[ki].split("(")[0][-52:]})
    e[r[mi]] = float(r[vi].replace(",", ""))
agg = collections.OrderedDict()
for e in per.values():
    key = (e["k"], int(e.get("launch__grid_size", 0)))
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += e.get("gpu__time_duration.sum", 0.0)
tot = sum(v for _, v in agg.values())
print("launches: %d   sum of kernel durations: %.1f us (cold-cache, serialised under ncu: compare SHARES)" % (len(per), tot / 1e3))
for (k, g), (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:48]:
    print("%-54s grid=%6d n=%3d total=%8.1f us  avg=%7.2f us  share=%5.1f%%" % (k, g, c, v / 1e3, v / c / 1e3, 100 * v / tot))
