"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import csv
import collections
import re
import sys

rows = []
with open(sys.argv[1]) as f:
    lines = [ln for ln in f if ln.startswith('"')]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "nsecond": 1e-3, "ms": 1e3, "msecond": 1e3}.get(unit, 1e-3)
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    rows.append((name, v * scale))
agg = collections.OrderedDict()
for n, t in rows:
    a = agg.setdefault(n, [0, 0.0])
    a[0] += 1
    a[1] += t
tot = sum(a[1] for a in agg.values())
print(f"{len(rows)} launches, {tot/1e3:.3f} ms total (cold-cache, serialised: compare SHARES)")
print(f"{'kernel':60s} {'launches':>8s} {'total us':>10s} {'avg us':>8s} {'share':>7s}")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:60]:60s} {c:8d} {t:10.1f} {t/c:8.2f} {100*t/tot:6.1f}%")
