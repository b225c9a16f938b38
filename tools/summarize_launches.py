"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: time per kernel name (+ template args)."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = defaultdict(lambda: [0, 0.0])
order = []
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    val = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = val / 1e3 if unit in ("ns", "nsecond") else (val if unit in ("us", "usecond") else val * 1e3)
    short = re.sub(r"\(.*$", "", name)
    grid = r.get("Grid Size", "")
    tot[short][0] += 1
    tot[short][1] += us
    order.append((short, us, grid))
total = sum(v[1] for v in tot.values())
print(f"total {total/1e3:.3f} ms over {sum(v[0] for v in tot.values())} launches")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{v[1]/1e3:9.3f} ms {100*v[1]/total:5.1f}%  n={v[0]:4d}  avg {v[1]/v[0]:8.1f} us  {k}")
if len(sys.argv) > 2:
    for s, us, g in order:
        print(f"{us:9.1f} us  {g:>14}  {s}")
