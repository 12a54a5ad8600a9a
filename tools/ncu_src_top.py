"""Top SASS instructions by warp-stall samples of an `ncu --page source --csv` export (one kernel)."""
import csv, sys
fn = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(fn)))
h = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[h]
body = []
for r in rows[h + 1:]:
    if r and r[0] in ("Kernel Name", "Address"):
        break                      # the export holds one section per captured launch: first one only
    if len(r) == len(hdr):
        body.append(r)
col = {c: i for i, c in enumerate(hdr)}
stalls = [c for c in hdr if c.startswith("stall_") and "Not Issued" not in c]
tot = sum(int(r[col["# Samples"]] or 0) for r in body)
agg = {s: sum(int(r[col[s]] or 0) for r in body) for s in stalls}
print("kernel:", rows[0][1][:100]); print("instructions:", len(body), "samples:", tot)
print("stall totals:", ", ".join("%s=%.1f%%" % (k[6:], 100.0 * v / max(tot, 1)) for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]))
order = sorted(range(len(body)), key=lambda i: -int(body[i][col["# Samples"]] or 0))[:n]
for i in sorted(order):
    r = body[i]; s = int(r[col["# Samples"]] or 0)
    top = sorted(((int(r[col[k]] or 0), k[6:]) for k in stalls), reverse=True)[:2]
    print("%5d %5.1f%%  %-70s ex=%s  %s" % (i, 100.0 * s / max(tot, 1), r[col["Source"]].strip()[:70], r[col["Instructions Executed"]],
                                        " ".join("%s:%d" % (k, v) for v, k in top if v)))
