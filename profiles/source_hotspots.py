"""Aggregate an `ncu --page source --csv --print-source sass,cuda` dump by CUDA source line (instructions + stall samples)."""
import csv, collections, sys
rows = list(csv.reader(open(sys.argv[1])))
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
agg = collections.defaultdict(lambda: [0.0, 0.0]); srcs = {}; cur = None; hdr = None
def num(x):
    try: return float(x)
    except (TypeError, ValueError): return 0.0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if r[0] == "Function Name" or hdr is None: continue
    try: ln = int(r[0])
    except ValueError: continue
    d = dict(zip(hdr, r))
    agg[(cur, ln)][0] += num(d.get("Instructions Executed")); agg[(cur, ln)][1] += num(d.get("# Samples")); srcs[(cur, ln)] = r[1]
ti = sum(v[0] for v in agg.values()) or 1; ts = sum(v[1] for v in agg.values()) or 1
print(f"total warp instructions {ti:.0f}, stall samples {ts:.0f}")
for (f, l), (i, s) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:topn]:
    print(f"{f}:{l:4d} inst {100*i/ti:5.1f}% samples {100*s/ts:5.1f}%  {srcs[(f, l)].strip()[:120]}")
