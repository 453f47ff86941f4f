"""A/B of QP settings in ONE process (same box, same clocks): python scripts/ab_settings.py "dict(active_set_start=20, active_set_step=20)" ..."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from centroidal_mpc_b200 import synthetic
from centroidal_mpc_b200.device import BatchSolver
conf = synthetic.load_conf(os.environ.get("AB_CONF", "solo12_trot"), N=100)
solver = BatchSolver(synthetic.make_batch(conf, int(os.environ.get("AB_BATCH", "4096"))))
variants = [None] + [eval(a) for a in sys.argv[1:]]
for rep in range(2):
    for qp in variants:
        for _ in range(2): solver.solve(conf.scp_params, qp)
        torch.cuda.synchronize(); ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); solver.solve(conf.scp_params, qp); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        st = solver.stats()
        print("%-60s min %.2f ms  median %.2f  nfac %.2f  admm %.1f" % (qp, min(ts), sorted(ts)[2], st["n_factor"].mean(), st["qp_iters"].mean()))
