"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel totals and shares of the
last <n> launches (one forward pass).  usage: launch_summary.py <csv> <launches per pass>"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
data = [(r[ki].split("(")[0][-58:], float(r[vi].replace(",", ""))) for r in rows[1:]]
n = int(sys.argv[2]) if len(sys.argv) > 2 else len(data)
last = data[-n:]
agg = collections.OrderedDict()
for k, v in last:
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(v for _, v in last)
print("launches: %d   sum of kernel durations: %.1f us (cold-cache, serialised under ncu: compare SHARES)" % (n, tot / 1e3))
for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-60s n=%3d total=%8.1f us  avg=%7.2f us  share=%5.1f%%" % (k, c, v / 1e3, v / c / 1e3, 100 * v / tot))
