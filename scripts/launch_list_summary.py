"""Aggregate an ncu launch list (``--metrics gpu__time_duration.sum --csv``): time and launch count per kernel.

    python scripts/launch_list_summary.py gpurun_out/x.csv [skip_first_n_launches]
"""
import csv, sys, collections, re

rows = []
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") == "gpu__time_duration.sum":
        v = float(r["Metric Value"].replace(",", ""))
        unit = r.get("Metric Unit", "ns")
        v *= {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "nsecond": 1e-6, "ms": 1.0, "msecond": 1.0}.get(unit, 1e-6)
        rows.append((r["Kernel Name"], v))
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = rows[skip:]
agg = collections.OrderedDict()
for k, v in rows:
    k = re.sub(r"<.*", "", k)[:70]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{len(rows)} launches, {tot:.3f} ms")
for k, (n, v) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{v:9.3f} ms {100 * v / tot:5.1f} %  x{n:<5d} {k}")
