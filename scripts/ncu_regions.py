"""Summarise an ncu report of cmpc_scp_kernel: headline metrics, then warp-stall samples and
executed instructions aggregated per source FUNCTION of csrc/*.cuh (needs -lineinfo and
--import-source on), then the hottest source lines.

  python scripts/ncu_regions.py gpurun_out/prof.ncu-rep [n_warps]
"""
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
nwarps = int(sys.argv[2]) if len(sys.argv) > 2 else 128


def ncu(*args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


raw = list(csv.reader(io.StringIO(ncu("--page", "raw", "--csv"))))
h, u, row = raw[0], raw[1], raw[2]
want = ["gpu__time_duration.sum", "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "sm__warps_active.avg.per_cycle_active",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "launch__shared_mem_per_block_dynamic",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "sass__inst_executed_local_loads", "sass__inst_executed_local_stores", "sass__inst_executed_shared_loads",
        "sass__inst_executed_shared_stores", "sass__inst_executed_global_loads", "sass__inst_executed_global_stores"]
for a, b, c in zip(h, u, row):
    if a in want or a.startswith("smsp__average_warps_issue_stalled") and "not_issued" not in a:
        try:
            if float(c) == 0.0:
                continue
        except ValueError:
            pass
        print("%-90s %-12s %s" % (a, b, c))

rows = list(csv.reader(io.StringIO(ncu("--page", "source", "--csv", "--print-source", "cuda,sass"))))
cur, agg, src = None, {}, {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1].split("/")[-1]
        continue
    if r[0] in ("Function Name", "Line No"):
        continue
    if r[0].isdigit() and r[2] == "-":
        agg[(cur, int(r[0]))] = (int(r[4] or 0), int(r[7] or 0))
        src[(cur, int(r[0]))] = r[1]
tot_s = sum(v[0] for v in agg.values()) or 1
tot_i = sum(v[1] for v in agg.values()) or 1

# map lines to functions by scanning the sources for definitions
import os
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
per = {}
for k, v in agg.items():
    name = "%s:%s" % (k[0], fn_of.get(k, "?"))
    s, i = per.get(name, (0, 0))
    per[name] = (s + v[0], i + v[1])
print("\nper function: stall samples, instructions executed")
for name, (s, i) in sorted(per.items(), key=lambda kv: -kv[1][0]):
    if s * 1000 < tot_s and i * 1000 < tot_i:
        continue
    print("  %-40s samples %5.1f%%   inst %5.1f%%  (%.3f M inst per warp)" % (name, 100.0 * s / tot_s, 100.0 * i / tot_i, i / nwarps / 1e6))
print("\nhottest lines")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][0])[:25]:
    print("  %-14s %4d  samples %5.2f%%  inst %5.2f%%  %s" % (k[0], k[1], 100.0 * v[0] / tot_s, 100.0 * v[1] / tot_i, src[k][:100]))
