"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
    python tools/summarise_launches.py gpurun_out/gn_launches.csv [-v]"""
import collections
import csv
import re
import sys

path = sys.argv[1]
with open(path) as f:
    lines = [ln for ln in f if not ln.startswith("==")]
agg = collections.OrderedDict()
seq = []
for row in csv.DictReader(lines):
    m = re.search(r"(sn_k_\w+|gn_\w+|solve_many|chi2_only|score_\w+|raster_\w+|\w+)\(", row["Kernel Name"] + "(")
    name = m.group(1) if m else row["Kernel Name"][:40]
    t = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    t = t / 1e3 if unit in ("ns", "nsecond") else t * 1e3 if unit in ("ms", "msecond") else t
    seq.append((name, t, row.get("Grid Size", "")))
    a = agg.setdefault(name, [0, 0.0, 0.0])
    a[0] += 1
    a[1] += t
    a[2] = max(a[2], t)
total = sum(a[1] for a in agg.values())
for k, a in agg.items():
    print("%-22s launches %5d  total %10.1f us  (%5.1f %%)  mean %8.2f  max %8.2f" %
          (k, a[0], a[1], 100 * a[1] / total, a[1] / a[0], a[2]))
print("total %.1f us over %d launches" % (total, len(seq)))
if "-v" in sys.argv:
    for name, t, grid in seq:
        print("%-22s %-16s %8.2f" % (name, grid.replace(" ", ""), t))
