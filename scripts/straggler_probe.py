"""How much of a wave is tail?  The headline batch against a batch of 4096 copies of ONE instance (every tile
does the same work, no instance waits for a neighbour's extra polish round) -- per-instance work taken from the
solver's own statistics.  python scripts/straggler_probe.py"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from centroidal_mpc_b200 import synthetic
from centroidal_mpc_b200.batch import ProblemBatch
from centroidal_mpc_b200.device import BatchSolver

conf = synthetic.load_conf("solo12_trot", N=100)
B = 4096
full = synthetic.make_batch(conf, B)


def timed(batch):
    s = BatchSolver(batch)
    for _ in range(3):
        s.solve(conf.scp_params)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); s.solve(conf.scp_params); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    st = s.stats()
    s.close()
    return float(np.median(ts)), st


ms, st = timed(full)
nf = st["n_factor"]
print(json.dumps({"batch": "mixed", "ms": ms, "factorisations_mean": float(nf.mean()),
                  "tile_max_mean": float(nf.reshape(-1, 4).max(1).mean()), "max": int(nf.max())}))
for target in sorted(set(int(v) for v in np.unique(nf))):
    b = int(np.nonzero(nf == target)[0][0])
    rep = lambda a: np.ascontiguousarray(np.broadcast_to(a[b:b + 1], a.shape))
    same = ProblemBatch.from_arrays(full.proto, rep(full.x_init), rep(full.x_final), rep(full.X_ref), rep(full.U_init))
    ms1, st1 = timed(same)
    print(json.dumps({"batch": "4096 copies of instance %d" % b, "ms": ms1, "factorisations": float(st1["n_factor"].mean()),
                      "multiplier_sweeps": float(st1["info"][:, 8].mean())}))
