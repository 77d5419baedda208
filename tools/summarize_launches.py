"""Summarise an `ncu --csv --metrics gpu__time_duration.sum` launch list by kernel name."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(float)
cnt = defaultdict(int)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"\(.*$", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1e-3)
    tot[name] += v * scale
    cnt[name] += 1
total = sum(tot.values())
print(f"total {total/1e3:.3f} ms over {sum(cnt.values())} launches")
print(f"{'kernel':70s} {'launches':>8s} {'ms':>10s} {'share':>7s} {'avg us':>9s}")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print(f"{k[:70]:70s} {cnt[k]:8d} {v/1e3:10.3f} {100*v/total:6.1f}% {v/cnt[k]:9.1f}")
