"""Multi-GPU path on CPU: shard arithmetic and the single gather, world_size 2 over gloo."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import emu_binding as E
from centroidal_mpc_b200 import parallel, synthetic


def test_shard_range_covers_batch():
    for B in (1, 7, 8, 4096, 4097):
        for world in (1, 2, 3, 8):
            spans = [parallel.shard_range(B, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == B
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, B, tmp):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    conf = synthetic.load_conf("solo12_trot", N=20)
    batch = synthetic.make_batch(conf, B)
    sub = parallel.shard_batch(batch, rank, world)
    out = E.solve_scp(sub, conf.scp_params)          # test stand-in for the CUDA solve of the shard
    local = dict(X=torch.from_numpy(out["X"]), U=torch.from_numpy(out["U"]),
                 ints=torch.from_numpy(np.stack([out["scp_iters"], out["status"], out["n_accepted"]])))
    full = parallel.gather_solutions(local, B, dist)
    if rank == 0:
        np.savez(tmp, X=full["X"].numpy(), U=full["U"].numpy(), ints=full["ints"].numpy())
    only0 = parallel.gather_solutions(local, B, dist, dst=0)
    assert (only0["X"] is None) == (rank != 0)
    dist.barrier()
    dist.destroy_process_group()


def test_sharded_solve_equals_single_rank(tmp_path):
    B, world = 5, 2                                   # ragged: shards of 3 and 2
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    tmp = str(tmp_path / "gathered.npz")
    mp.spawn(_worker, args=(world, port, B, tmp), nprocs=world, join=True)
    got = np.load(tmp)
    conf = synthetic.load_conf("solo12_trot", N=20)
    ref = E.solve_scp(synthetic.make_batch(conf, B), conf.scp_params)
    np.testing.assert_array_equal(got["X"], ref["X"])
    np.testing.assert_array_equal(got["U"], ref["U"])
    np.testing.assert_array_equal(got["ints"][0], ref["scp_iters"])
