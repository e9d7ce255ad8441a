"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list of bench.py into one step's table:
per (kernel, grid) totals.  Usage: python tools/launch_table.py launches.csv [steps_in_capture]"""
import csv, sys, collections, re

path = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("u3d::(anonymous namespace)::", "").replace("void ", "")
    name = re.sub(r"at::native::", "", name)[:70]
    rows.append((name, r["Grid Size"], us))
agg = collections.OrderedDict()
for name, grid, us in rows:
    k = (name, grid)
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1
    a[1] += us
byname = collections.defaultdict(lambda: [0, 0.0])
for (name, grid), (n, us) in agg.items():
    byname[name][0] += n
    byname[name][1] += us
tot = sum(v[1] for v in byname.values())
print(f"total {tot / steps / 1e3:.2f} ms per step over {len(rows) / steps:.0f} launches per step")
for name, (n, us) in sorted(byname.items(), key=lambda kv: -kv[1][1]):
    print(f"{us / steps / 1e3:9.3f} ms  x{n / steps:7.1f}  {name}")
if "--detail" in sys.argv:
    for (name, grid), (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:80]:
        print(f"{us / steps / 1e3:9.3f} ms  x{n / steps:6.1f}  avg {us / n:8.1f} us  grid {grid:>18s}  {name}")
