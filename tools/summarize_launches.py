"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name."""
import csv
import collections
import re
import sys

path = sys.argv[1]
rows = []
with open(path, newline="") as fh:
    lines = [l for l in fh if not l.startswith("==")]
rd = csv.DictReader(lines)
tot = collections.defaultdict(float)
cnt = collections.Counter()
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = r["Kernel Name"]
    name = re.sub(r"<.*", "", name)
    name = re.sub(r"\(.*", "", name)
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    tot[name] += us
    cnt[name] += 1
total = sum(tot.values())
print("kernel,launches,total_us,share,avg_us")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print("%s,%d,%.1f,%.3f,%.2f" % (k, cnt[k], v, v / total, v / cnt[k]))
print("TOTAL,%d,%.1f,1.000," % (sum(cnt.values()), total))
