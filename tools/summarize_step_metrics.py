"""Per-kernel summary of an ncu multi-metric CSV (time, DRAM bytes, tensor activity)."""
import csv, collections, re, sys
lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
rd = csv.DictReader(lines)
per = collections.defaultdict(lambda: collections.defaultdict(float))
cnt = collections.Counter()
seen = set()
for r in rd:
    name = re.sub(r"\(.*", "", re.sub(r"<.*", "", r["Kernel Name"]))
    key = (r["ID"], name)
    v = float(r["Metric Value"].replace(",", "")) if r["Metric Value"] not in ("", "n/a") else 0.0
    m, u = r["Metric Name"], r["Metric Unit"]
    if m == "gpu__time_duration.sum":
        v = v / 1e3 if u.startswith("n") else (v if u.startswith("u") else v * 1e3)
        if key not in seen:
            cnt[name] += 1; seen.add(key)
    if m.startswith("dram__bytes") or m == "lts__t_bytes.sum":
        mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
        v *= mult
    per[name][m] += v
tot = sum(d["gpu__time_duration.sum"] for d in per.values())
print("kernel,launches,total_us,share,avg_us,dram_read_MB_per_launch,dram_write_MB_per_launch,l2_MB_per_launch,tensor_active_pct_avg")
for k, d in sorted(per.items(), key=lambda kv: -kv[1]["gpu__time_duration.sum"])[:25]:
    n = cnt[k]
    print("%s,%d,%.1f,%.3f,%.2f,%.2f,%.2f,%.1f,%.1f" % (k, n, d["gpu__time_duration.sum"], d["gpu__time_duration.sum"] / tot,
          d["gpu__time_duration.sum"] / n, d["dram__bytes_read.sum"] / n / 1e6, d["dram__bytes_write.sum"] / n / 1e6,
          d["lts__t_bytes.sum"] / n / 1e6, d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"] / n))
print("TOTAL,%d,%.1f" % (sum(cnt.values()), tot))
