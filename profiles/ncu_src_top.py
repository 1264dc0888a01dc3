"""Summarise `ncu --page source --csv` output: top stall sites of the first kernel in the report."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
si, ci = hdr.index("# Samples"), hdr.index("Source")
stalls = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
data = []
for r in rows[2:]:
    if r and r[0] == "Address":
        break
    try:
        n = int(r[si])
    except Exception:
        continue
    top = sorted(((int(r[i] or 0), hdr[i]) for i in stalls), reverse=True)[:2]
    data.append((n, r[ci][:90], top))
tot = sum(d[0] for d in data)
print("total samples", tot)
for d in sorted(data, reverse=True)[: int(sys.argv[2]) if len(sys.argv) > 2 else 20]:
    print("%5.1f%%  %-90s %s" % (100.0 * d[0] / tot, d[1], d[2]))
