"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (share of total)."""
import csv, re, sys
tot, n = {}, {}
rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
for r in csv.DictReader(rows):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k = re.sub(r"\(.*", "", r["Kernel Name"])
    v = float(r["Metric Value"].replace(",", ""))
    u = r["Metric Unit"]
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(u, 1e-6)
    tot[k] = tot.get(k, 0) + v
    n[k] = n.get(k, 0) + 1
s = sum(tot.values())
print("# %d launches, %.3f ms total (ncu per-launch times: cold cache, serialised -> compare shares)" % (sum(n.values()), s))
for k in sorted(tot, key=tot.get, reverse=True):
    print("%-64s n=%4d %9.3f ms %5.1f%%" % (k[:64], n[k], tot[k], 100 * tot[k] / s))
