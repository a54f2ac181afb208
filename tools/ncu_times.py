"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` log: total/mean duration per kernel name."""
import collections
import csv
import re
import sys

rows = list(csv.reader(l for l in open(sys.argv[1], errors="replace") if l.startswith('"')))
hdr = rows[0]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
ig = hdr.index("Grid Size") if "Grid Size" in hdr else None
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= iv:
        continue
    name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("<unnamed>::", "")
    key = (name, (r[ig] if ig is not None else "") if "--by-grid" in sys.argv else "")
    a = agg.setdefault(key, [0, 0.0])
    a[0] += 1
    a[1] += float(r[iv].replace(",", ""))
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':48s} {'grid':>16s} {'n':>4s} {'mean_us':>10s} {'total_us':>10s} {'share':>6s}")
for (name, grid), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:48]:48s} {grid:>16s} {n:4d} {t / n / 1e3:10.1f} {t / 1e3:10.1f} {100 * t / tot:5.1f}%")
