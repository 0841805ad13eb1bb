"""One simulation step out of an `ncu --metrics gpu__time_duration.sum --csv` launch list:
kernels between two launches of the step's first user node (argv[2], default movementSystem)."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
first = sys.argv[2] if len(sys.argv) > 2 else "movementSystem"
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
H = rows[hdr]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
seq = []
for r in rows[hdr + 1:]:
    if len(r) > vi:
        v = float(r[vi].replace(",", ""))
        if r[ui] == "ns":
            v /= 1000
        seq.append((r[ki][:96], v))
idx = [i for i, (k, v) in enumerate(seq) if first in k]
a, b = idx[0], idx[1]
agg = collections.OrderedDict()
tot = 0
for k, v in seq[a:b]:
    agg.setdefault(k, [0, 0])
    agg[k][0] += v
    agg[k][1] += 1
    tot += v
for k, (v, n) in agg.items():
    print("%-98s x%-2d %8.1f us" % (k, n, v))
print("sum %.1f us, %d launches (cold-cache, serialised: compare shares)" % (tot, b - a))
