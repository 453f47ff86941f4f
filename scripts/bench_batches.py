"""Batch sweep of one workload on one GPU (informational): python scripts/bench_batches.py [workload] [B ...]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from bench_configs import run

if __name__ == "__main__":
    name = sys.argv[1] if len(sys.argv) > 1 else "solo12_trot"
    for B in [int(x) for x in sys.argv[2:]] or (256, 1024, 4096, 8192, 16384, 65536):
        run(name, 100, B, reps=3)
