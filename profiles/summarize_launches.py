"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel name count / total / mean / share.
  python profiles/summarize_launches.py gpurun_out/x_launches.csv [first_id last_id]"""
import csv
import re
import sys
from collections import OrderedDict

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1000.0 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1000.0)
    rows.append((int(r["ID"]), re.sub(r"\(.*", "", r["Kernel Name"]), us, r.get("Grid Size", ""), r.get("Block Size", "")))
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 9
rows = [r for r in rows if lo <= r[0] <= hi]
agg = OrderedDict()
for _, name, us, g, b in rows:
    a = agg.setdefault(name, [0, 0.0, 1e30, 0.0])
    a[0] += 1; a[1] += us; a[2] = min(a[2], us); a[3] = max(a[3], us)
tot = sum(a[1] for a in agg.values())
print(f"{len(rows)} launches, {tot:.1f} us total (ids {rows[0][0]}..{rows[-1][0]})")
print(f"{'kernel':60s} {'n':>5s} {'total us':>10s} {'mean':>8s} {'min':>8s} {'max':>8s} {'share':>6s}")
for name, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:60]:60s} {a[0]:5d} {a[1]:10.1f} {a[1]/a[0]:8.2f} {a[2]:8.2f} {a[3]:8.2f} {100*a[1]/tot:5.1f}%")
