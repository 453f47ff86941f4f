"""Receding-horizon loop on one GPU: ms per tick (B MPC loops advance together), warm-started vs cold.
  python scripts/bench_mpc.py [workload] [B] [H] [ticks]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from centroidal_mpc_b200.mpc import RecedingHorizonMPC

name = sys.argv[1] if len(sys.argv) > 1 else "solo12_trot"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
H = int(sys.argv[3]) if len(sys.argv) > 3 else 100
ticks = int(sys.argv[4]) if len(sys.argv) > 4 else 24
for warm in (True, False):
    mpc = RecedingHorizonMPC(name, B, H, warm=warm)
    ts, qp, nf, pm = [], [], [], []
    for t in range(ticks):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); mpc.step(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
        st = mpc.stats()
        qp.append(float(st["qp_iters"].mean())); nf.append(float(st["n_factor"].mean())); pm.append(float(st["info"][:, 8].mean()))
    print(json.dumps({"workload": name, "batch": B, "horizon": H, "ticks": ticks, "warm_start": warm, "first_tick_ms": ts[0],
                      "ms_per_tick_p50": float(np.median(ts[1:])), "ticks_x_instances_per_s": B / float(np.median(ts[1:])) * 1e3,
                      "admm_iters_mean": float(np.mean(qp[1:])), "factorisations_mean": float(np.mean(nf[1:])),
                      "multiplier_sweeps_mean": float(np.mean(pm[1:])), "failed_instances": int(mpc.failed.sum())}), flush=True)
    mpc.close()
