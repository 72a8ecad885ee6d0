"""Hottest SASS instructions (warp-stall samples) of one launch in an .ncu-rep captured with --import-source on:
    python tools/ncu_hot.py x.ncu-rep [launch_index] [top]"""
import csv, subprocess, sys
rep = sys.argv[1]
idx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--launch-skip", str(idx), "--launch-count", "1"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
print(rows[0][:2])
h = rows[hi]
isrc, isamp, iex = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
data = []
for n, r in enumerate(rows[hi + 1:]):
    try:
        data.append((int(r[isamp] or 0), int(r[iex] or 0), r[isrc].strip(), n))
    except Exception:
        pass
tot = sum(d[0] for d in data)
print("total samples", tot, "instructions", len(data))
for s, e, src, n in sorted(data, reverse=True)[:top]:
    print("%6d %5.1f%% ex=%9d  #%5d  %s" % (s, 100.0 * s / tot, e, n, src[:100]))
