"""Timeline of one cmpc_solve_scp_host call on the headline workload (run on the GPU box): kernel and
copy activities per stream from CUPTI (torch.profiler), to check that the chunks' copies overlap the solves."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import ProfilerActivity, profile
from centroidal_mpc_b200 import synthetic
from centroidal_mpc_b200.device import BatchSolver

conf = synthetic.load_conf("solo12_trot", N=100)
B = int(os.environ.get("TL_BATCH", "4096"))
batch = synthetic.make_batch(conf, B)
keep = []
for name in ("x_init", "x_final", "X_ref", "U_init", "contact_pos", "contact_active"):
    t = torch.from_numpy(getattr(batch, name)).pin_memory(); keep.append(t); setattr(batch, name, t.numpy())
solver = BatchSolver(batch)
out = dict(X=torch.empty((B, 101, 9), dtype=torch.float64).pin_memory(), U=torch.empty((B, 100, 12), dtype=torch.float64).pin_memory(),
           scp_iters=torch.empty(B, dtype=torch.int32).pin_memory(), status=torch.empty(B, dtype=torch.int32).pin_memory(),
           n_accepted=torch.empty(B, dtype=torch.int32).pin_memory())
out_np = {k: v.numpy() for k, v in out.items()}
for _ in range(3):
    solver.solve_host(conf.scp_params, out=out_np)
ts = []
for _ in range(5):
    torch.cuda.synchronize(); t0 = time.perf_counter(); solver.solve_host(conf.scp_params, out=out_np); ts.append((time.perf_counter() - t0) * 1e3)
print("solve_host wall ms:", " ".join("%.2f" % t for t in ts))
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    solver.solve_host(conf.scp_params, out=out_np)
    torch.cuda.synchronize()
evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
evs.sort(key=lambda e: e.time_range.start)
t0 = evs[0].time_range.start
for e in evs:
    if e.time_range.end - e.time_range.start < 20 and "scp" not in e.name:
        continue
    print("%9.3f ms  +%8.3f ms  %s" % ((e.time_range.start - t0) / 1e3, (e.time_range.end - e.time_range.start) / 1e3, e.name[:60]))
