"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: time share per kernel.
usage: python tools/summarize_launches.py gpurun_out/launches.csv > profiles/rNN_launches.md"""
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
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    name = re.sub(r"<.*", "", name).replace("amb::", "")
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    ns = v * {"ns": 1, "us": 1e3, "usecond": 1e3, "nsecond": 1, "ms": 1e6, "msecond": 1e6, "second": 1e9, "s": 1e9}.get(unit, 1)
    tot[name] += ns
    cnt[name] += 1
total = sum(tot.values())
print(f"# ncu launch list summary ({path})\n")
print(f"{sum(cnt.values())} launches, {total / 1e9:.3f} s of kernel time (cold-cache, serialised under ncu: compare SHARES)\n")
print("| kernel | launches | total ms | share | avg us |")
print("|---|---:|---:|---:|---:|")
for k in sorted(tot, key=tot.get, reverse=True):
    print(f"| {k} | {cnt[k]} | {tot[k] / 1e6:.2f} | {100 * tot[k] / total:.2f} % | {tot[k] / cnt[k] / 1e3:.1f} |")
