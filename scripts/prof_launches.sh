#!/bin/bash
# per-launch device times of 2 RL iterations (cold-cache, serialised: compare SHARES, not absolutes)
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python scripts/prof_run.py > gpurun_out/prof_launches.log 2>&1
python - <<'PY'
import csv, collections
rows = []
with open("gpurun_out/launches.csv") as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.DictReader(lines)
agg = collections.OrderedDict()
for row in r:
    if row.get("Metric Name") != "gpu__time_duration.sum": continue
    name = row["Kernel Name"].split("(")[0][:70]
    v = float(row["Metric Value"].replace(",", "")); unit = row["Metric Unit"]
    if unit == "ns": v /= 1e3
    elif unit == "ms": v *= 1e3
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f"{'kernel':70s} {'n':>5s} {'total us':>10s} {'avg us':>9s} {'share':>6s}")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:70s} {n:5d} {t:10.1f} {t/n:9.1f} {100*t/tot:5.1f}%")
PY
