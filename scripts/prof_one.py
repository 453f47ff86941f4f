"""Three batched solves of the benchmark workload (for ncu: profile the last launch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from centroidal_mpc_b200 import synthetic
from centroidal_mpc_b200.device import BatchSolver

name = sys.argv[1] if len(sys.argv) > 1 else "solo12_trot"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
conf = synthetic.load_conf(name, N=100)
STOCH = len(sys.argv) > 3 and sys.argv[3] == "stoch"   # stochastic mode (general friction path)
solver = BatchSolver(synthetic.make_batch(conf, B, stochastic=STOCH))
for _ in range(3):
    solver.solve(conf.scp_params)
torch.cuda.synchronize()
print("status", solver.results()["status"].sum())
