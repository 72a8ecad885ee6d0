"""profiles/traffic.json: DRAM bytes (read + write) per launch of the main kernels, from `ncu --set full` captures:
    python tools/make_traffic.py <tag> a.ncu-rep b.ncu-rep c.ncu-rep:key=regex ...   (tag = the session of the captures)
A plain report is matched against the kernel-name table below; `report:key=regex` files every launch whose kernel name
matches `regex` under `key` (the density-grid program and the delta-streaming skin warp share their kernel templates
with the training kernels).  Entries of earlier sessions that are not re-measured stay in the file."""
import csv, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
_T = r"\(?(?:int)?\)?"   # ncu prints template arguments as (int)128


def _chain(width, prog):
    return r"chain_kernel<%s%d, %s\d+, %s\d+, %s\d+, %s%d[,>]" % (_T, width, _T, _T, _T, _T, prog)


NAMES = [(_chain(128, 0), "moda_chain_trunk_fwd"), (_chain(128, 1), "moda_chain_trunk_bwd"),
         (_chain(64, 0), "moda_chain_skin_fwd"), (_chain(64, 1), "moda_chain_skin_bwd"),
         (r"tc_wgrad_kernel<%s256, %s256>" % (_T, _T), "moda_tc_wgrad<256,256>"),
         (r"tc_wgrad_multi_kernel<%s256, %s256>" % (_T, _T), "moda_tc_wgrad_multi<256,256>"),
         (r"tc_wgrad_multi_kernel<%s64, %s64>" % (_T, _T), "moda_tc_wgrad_multi<64,64>"),
         (r"tc_wgrad_multi_kernel<%s256, %s64>" % (_T, _T), "moda_tc_wgrad_multi<256,64>"),
         (r"tc_wgrad_kernel<%s128, %s256>" % (_T, _T), "moda_tc_wgrad<128,256>"),
         (r"skin_warp_fwd", "moda_skin_warp_fwd"), (r"skin_warp_bwd", "moda_skin_warp_bwd")]
SCALE = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tag, reps = sys.argv[1], sys.argv[2:]
acc = {}
for rep in reps:
    forced = None
    if ":" in rep and "=" in rep.split(":", 1)[1]:
        rep, spec = rep.split(":", 1)
        forced = spec.split("=", 1)
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    h, units = rows[0], rows[1]
    u = dict(zip(h, units))
    for r in rows[2:]:
        d = dict(zip(h, r))
        name = d.get("Kernel Name", "")
        if forced is not None:
            entry = forced[0] if re.search(forced[1], name) else None
        else:
            entry = next((e for pat, e in NAMES if re.search(pat, name)), None)
        if entry is None:
            continue
        b = sum(float(d[m]) * SCALE[u[m]] for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        t = float(d["gpu__time_duration.sum"]) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(u["gpu__time_duration.sum"], 1.0)
        a = acc.setdefault(entry, {"kernel": name.split("(")[0], "bytes": [], "us": [], "report": os.path.basename(rep)})
        a["bytes"].append(b)
        a["us"].append(t)
path = os.path.join(ROOT, "profiles", "traffic.json")
res = json.load(open(path)) if os.path.exists(path) else {"kernels": {}}
res["source"] = "ncu --set full --clock-control none, per launch, cold cache; session per entry"
for e, a in acc.items():
    res["kernels"][e] = {"kernel": a["kernel"], "dram_bytes_per_launch": sum(a["bytes"]) / len(a["bytes"]),
                         "launches_captured": len(a["bytes"]), "us_per_launch_under_ncu": sum(a["us"]) / len(a["us"]),
                         "report": a["report"], "session": tag}
json.dump(res, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
print(json.dumps(res, indent=1))
