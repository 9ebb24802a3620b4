"""Summarise an ncu launch list (--metrics gpu__time_duration.sum --csv): per kernel count, mean
duration and share of the summed kernel time.  usage: launch_summary.py launches.csv"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = None
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows:
    if hdr is None:
        if "Kernel Name" in r:
            hdr = {h: i for i, h in enumerate(r)}
        continue
    if len(r) < len(hdr) or r[hdr["Metric Name"]] != "gpu__time_duration.sum":
        continue
    n = re.sub(r"\(.*", "", r[hdr["Kernel Name"]])
    v = float(r[hdr["Metric Value"]].replace(",", ""))
    u = r[hdr["Metric Unit"]]
    v = v / 1e3 if u == "ns" else v * 1e3 if u == "ms" else v
    agg[n][0] += 1
    agg[n][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':58s} {'n':>5s} {'mean us':>9s} {'share':>7s}")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n:58s} {c:5d} {t / c:9.1f} {100 * t / tot:6.2f}%")
