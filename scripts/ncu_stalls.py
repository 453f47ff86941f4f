"""Per source function of csrc/*: executed instructions and warp-stall samples by reason (ncu --set full, --import-source on).
   python scripts/ncu_stalls.py gpurun_out/prof.ncu-rep"""
import csv, io, os, re, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hdr = None
cur = None
agg = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if r[0].isdigit() and r[2] == "-" and hdr:
        d = dict(zip(hdr[4:], r[4:]))
        agg[(cur, int(r[0]))] = d
root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "centroidal_mpc_b200", "csrc")
fn_of = {}
for f in set(k[0] for k in agg):
    path = os.path.join(root, f)
    if not os.path.exists(path):
        continue
    name = "(top)"
    for n, line in enumerate(open(path), 1):
        m = re.match(r"^(?:CMPC_HD|CMPC_FN|__device__|__global__|template|inline|static)?.*?\b([A-Za-z_][A-Za-z_0-9]*)\(.*[,{)]\s*$", line)
        if m and not line.startswith(" ") and not line.startswith("//") and not line.startswith("#"):
            name = m.group(1)
        fn_of[(f, n)] = name
keys = ["stall_no_inst", "stall_wait", "stall_long_sb", "stall_short_sb", "stall_branch_resolving", "stall_selected", "stall_math", "stall_barrier"]
per = {}
for k, d in agg.items():
    name = fn_of.get(k, k[0])
    p = per.setdefault(name, dict(inst=0, samples=0, **{x: 0 for x in keys}))
    p["inst"] += int(d.get("Instructions Executed") or 0)
    p["samples"] += int(d.get("# Samples") or 0)
    for x in keys:
        p[x] += int(d.get(x) or 0)
tot = sum(p["samples"] for p in per.values()) or 1
toti = sum(p["inst"] for p in per.values()) or 1
print("%-24s %7s %7s | %s" % ("function", "inst%", "smpl%", " ".join("%9s" % x.replace("stall_", "")[:9] for x in keys)))
for name, p in sorted(per.items(), key=lambda kv: -kv[1]["samples"]):
    if p["samples"] * 200 < tot:
        continue
    print("%-24s %7.1f %7.1f | %s" % (name[:24], 100.0 * p["inst"] / toti, 100.0 * p["samples"] / tot,
                                     " ".join("%9.1f" % (100.0 * p[x] / max(p["samples"], 1)) for x in keys)))
allp = {x: sum(p[x] for p in per.values()) for x in keys}
print("%-24s %7s %7s | %s" % ("ALL", "", "", " ".join("%9.1f" % (100.0 * allp[x] / tot) for x in keys)))
