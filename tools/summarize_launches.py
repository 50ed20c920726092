"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into a markdown table.
Usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/<name>.md"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5 and r[0].isdigit()]
tot, cnt = defaultdict(float), defaultdict(int)
for r in rows:
    name = r[4].split("(")[0].replace("<unnamed>::", "").replace("void ", "")[:70]
    v, unit = float(r[-1].replace(",", "")), r[-2]
    v *= {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}[unit]
    tot[name] += v
    cnt[name] += 1
s = sum(tot.values())
print("| kernel | launches | total ms | share |")
print("|---|---:|---:|---:|")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    if v / s < 5e-4:
        continue
    print("| `%s` | %d | %.3f | %.1f%% |" % (k, cnt[k], v, 100 * v / s))
