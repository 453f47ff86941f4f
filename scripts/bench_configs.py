"""The other BASELINE.json configurations and a batch sweep on one GPU (informational; the
headline line comes from bench.py).  Prints one JSON line per run.
  python scripts/bench_configs.py > gpurun_out/configs.jsonl"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from centroidal_mpc_b200 import synthetic
from centroidal_mpc_b200.device import BatchSolver


def run(name, N, B, mode="B", reps=5, stochastic=False):
    conf = synthetic.load_conf(name, N=N)
    solver = BatchSolver(synthetic.make_batch(conf, B, mode=mode, stochastic=stochastic))
    for _ in range(3):
        solver.solve(conf.scp_params)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); solver.solve(conf.scp_params); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    res, st = solver.results(), solver.stats()
    ms = float(np.median(ts))
    print(json.dumps({"workload": name, "N": N, "batch": B, "mode": mode, "stochastic": stochastic, "ms_p50": ms, "solves_per_s": B / ms * 1e3,
                      "failed": int((res["status"] != 0).sum()), "accepted": int((res["n_accepted"] > 0).sum()),
                      "scp_iters_mean": float(res["scp_iters"].mean()), "admm_iters_mean": float(st["qp_iters"].mean()),
                      "admm_iters_max": int(st["qp_iters"].max()), "factorisations_mean": float(st["n_factor"].mean())}), flush=True)
    solver.close()


if __name__ == "__main__":
    run("solo12_trot", 100, 1)                      # single-instance latency
    run("solo12_pace", 100, 1024, mode="A")         # config 2: perturbed initial states
    run("solo12_bound", 100, 4096)                  # config 3
    run("bolt", 100, 1024)                          # config 4: 8192 over 8 GPUs = 1024 per GPU
    run("bolt", 100, 8192)                          # ... and the whole batch on one GPU
    run("solo12_trot", 100, 4096, stochastic=True)  # stochastic mode: friction rows with chance-constraint back-offs
    run("solo12_bound", 100, 4096, stochastic=True)
    for B in (256, 1024, 4096, 16384, 65536):       # the batch sweep on the headline problem
        run("solo12_trot", 100, B, reps=3)
    for B in (256, 1024, 4096, 16384, 65536):       # config 5: talos (CoP / wrench contact model), batch sweep
        run("talos", 100, B, reps=3)
