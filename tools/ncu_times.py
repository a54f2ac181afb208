"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` log:
launches, mean / total duration and mean DRAM bytes per kernel name."""
import collections
import csv
import re
import sys

rows = list(csv.reader(l for l in open(sys.argv[1], errors="replace") if l.startswith('"')))
hdr = rows[0]
ik, iv, im, iu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name"), hdr.index("Metric Unit")
ig = hdr.index("Grid Size") if "Grid Size" in hdr else None
SCALE = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6,
         "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) <= iv:
        continue
    name = re.sub(r"\(.*", "", r[ik]).replace("void ", "").replace("<unnamed>::", "")
    key = (name, (r[ig] if ig is not None else "") if "--by-grid" in sys.argv else "")
    a = agg.setdefault(key, {"n": 0, "us": 0.0, "rd": 0.0, "wr": 0.0})
    v = float(r[iv].replace(",", "")) * SCALE.get(r[iu], 1.0)
    if r[im].startswith("gpu__time_duration"):
        a["n"] += 1
        a["us"] += v
    elif r[im].startswith("dram__bytes_read"):
        a["rd"] += v
    elif r[im].startswith("dram__bytes_write"):
        a["wr"] += v
tot = sum(a["us"] for a in agg.values())
print(f"{'kernel':48s} {'grid':>12s} {'n':>4s} {'mean_us':>10s} {'total_us':>10s} {'share':>6s} {'rd_MB':>9s} {'wr_MB':>9s}")
for (name, grid), a in sorted(agg.items(), key=lambda kv: -kv[1]["us"]):
    n = max(a["n"], 1)
    print(f"{name[:48]:48s} {grid:>12s} {a['n']:4d} {a['us'] / n:10.1f} {a['us']:10.1f} {100 * a['us'] / tot:5.1f}% "
          f"{a['rd'] / n / 1e6:9.1f} {a['wr'] / n / 1e6:9.1f}")
