"""A/B timing of library variants on the headline workload (run on the GPU box):
   python scripts/ab.py variants/libcmpc_a.so variants/libcmpc_b.so ...   (paths relative to csrc)
Each variant runs in its own process: 3 warm-up + 5 timed batched solves, prints ms per solve."""
import os, subprocess, sys
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
child = r'''
import os, sys, time
sys.path.insert(0, %r)
import torch
from centroidal_mpc_b200 import synthetic
from centroidal_mpc_b200.device import BatchSolver
conf = synthetic.load_conf("solo12_trot", N=100)
B = int(os.environ.get("AB_BATCH", "4096"))
solver = BatchSolver(synthetic.make_batch(conf, B))
for _ in range(3): solver.solve(conf.scp_params)
torch.cuda.synchronize()
ts = []
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); solver.solve(conf.scp_params); e1.record(); torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
r = solver.results()
print("%%-34s B %%6d ms %%s  min %%.2f  %%.0f solves/s  failed %%d" %% (os.environ.get("CMPC_B200_LIB", "default")[-34:], B, " ".join("%%.2f" %% t for t in ts), min(ts), B / min(ts) * 1e3, int((r["status"] != 0).sum())))
''' % root
for lib in sys.argv[1:] or [""]:
    env = dict(os.environ)
    if lib:
        env["CMPC_B200_LIB"] = os.path.join(root, "centroidal_mpc_b200", "csrc", lib)
    subprocess.run([sys.executable, "-c", child], env=env, timeout=300)
