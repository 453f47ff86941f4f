"""Turn the files a GPU run left under gpurun_out/ into the tracked summaries under profiles/:
  python scripts/make_profiles.py TAG
reads  gpurun_out/{bench_r2_TAG.json, launches_TAG.csv, prof_TAG.ncu-rep, configs_r2.jsonl}
writes profiles/r2_TAG_{bench.json, launches.csv, scp_kernel.txt, configs.jsonl} and profiles/r2_traffic.json (stamped with the
build id of the library in the tree: run it right after the GPU call, before rebuilding)"""
import collections, csv, json, os, re, subprocess, sys
tag = sys.argv[1]
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
go, pr = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles")
line = open(os.path.join(go, "bench_r2_%s.json" % tag)).read().strip().splitlines()[-1]
open(os.path.join(pr, "r2_%s_bench.json" % tag), "w").write(line + "\n")
rows = [r for r in csv.reader(open(os.path.join(go, "launches_%s.csv" % tag))) if r and not r[0].startswith("==")]
hdr = rows[0]; ix = {n: i for i, n in enumerate(hdr)}
agg = collections.OrderedDict()
for r in rows[1:]:
    if len(r) < len(hdr):
        continue
    v, unit = float(r[ix["Metric Value"]]), r[ix["Metric Unit"]]
    ms = v / 1e6 if unit.startswith("ns") else (v / 1e3 if unit.startswith("us") else v)
    a = agg.setdefault(r[ix["Kernel Name"]], [0, 0.0]); a[0] += 1; a[1] += ms
tot = sum(a[1] for a in agg.values())
out = ["# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 2 --warmup 3 --no-cpu-baseline",
       "# (per-launch times under ncu are serialised and cold-cache; the SHARE of the step is what counts;",
       "#  cmpc_dfma_kernel is the FP64 peak micro-benchmark bench.py runs outside the timed region)",
       "kernel,launches,total_ms,ms_per_launch,share"]
for n, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append("%s,%d,%.3f,%.3f,%.4f" % (n[:90].replace(",", ";"), c, ms, ms / c, ms / tot))
open(os.path.join(pr, "r2_%s_launches.csv" % tag), "w").write("\n".join(out) + "\n")
t = subprocess.run([sys.executable, os.path.join(root, "scripts", "ncu_regions.py"), os.path.join(go, "prof_%s.ncu-rep" % tag), "1024"],
                   capture_output=True, text=True).stdout
hdr_txt = """# cmpc_scp_kernel, solo12 trot N=100, batch 4096 (one launch), B200
# ncu --set full --clock-control none --import-source on -k regex:cmpc_scp -s 2 -c 1 python scripts/prof_one.py
# summarised by scripts/ncu_regions.py (headline metrics; stall samples / executed instructions per source function; hottest lines)
# One-warp CTAs: a warp is a tile of 4 instances x 8 lanes; 1024 CTAs, 7 per SM; 'per warp' divides by the 1024 CTAs.
"""
t += "\n# warp-stall samples by reason per source function (scripts/ncu_stalls.py)\n" + subprocess.run(
    [sys.executable, os.path.join(root, "scripts", "ncu_stalls.py"), os.path.join(go, "prof_%s.ncu-rep" % tag)], capture_output=True, text=True).stdout
open(os.path.join(pr, "r2_%s_scp_kernel.txt" % tag), "w").write(hdr_txt + t)
rd = float(re.search(r"dram__bytes_read.sum\s+Gbyte\s+([\d.]+)", t).group(1))
wr = float(re.search(r"dram__bytes_write.sum\s+Gbyte\s+([\d.]+)", t).group(1))
ms = float(re.search(r"gpu__time_duration.sum\s+ms\s+([\d.]+)", t).group(1))
import ctypes
so = ctypes.CDLL(os.path.join(root, "centroidal_mpc_b200", "csrc", "libcmpc_b200.so"))
so.cmpc_build_id.restype = ctypes.c_char_p
json.dump({"build_id": so.cmpc_build_id().decode(), "workload": "solo12_trot N=100 B=4096", "dram_bytes_per_launch": (rd + wr) * 1e9, "dram_bytes_read": rd * 1e9, "dram_bytes_write": wr * 1e9, "kernel_ms_under_ncu": ms,
           "source": "ncu --set full --clock-control none -k regex:cmpc_scp -s 2 -c 1 python scripts/prof_one.py (solo12_trot N=100, batch 4096); profiles/r2_%s_scp_kernel.txt" % tag},
          open(os.path.join(pr, "r2_traffic.json"), "w"), indent=1)
cfg = os.path.join(go, "configs_r2.jsonl")
if os.path.exists(cfg):
    open(os.path.join(pr, "r2_%s_configs.jsonl" % tag), "w").write(open(cfg).read())
print("\n".join(out[3:7])); print("traffic GB", rd + wr, "kernel ms", ms)
