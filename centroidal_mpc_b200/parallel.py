"""Multi-GPU: MPC instances are independent, so the batch is split contiguously across ranks
(one process per GPU), every rank solves its shard with no collective on the solve path, and
one gather collects the solutions.  ``torch.distributed`` is plumbing:
NCCL on GPUs, gloo in the CPU tests."""
import numpy as np


def shard_range(B, rank, world):
    """Contiguous split; the first B % world ranks get one extra instance."""
    base, extra = divmod(int(B), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(batch, rank, world):
    """ProblemBatch holding this rank's instances (views of the parent's arrays)."""
    import copy
    lo, hi = shard_range(batch.B, rank, world)
    sub = copy.copy(batch)
    sub.B = hi - lo
    for name in ("x_init", "x_final", "X_ref", "U_init"):
        setattr(sub, name, np.ascontiguousarray(getattr(batch, name)[lo:hi]))
    if not batch.shared_plan:
        for name in ("contact_pos", "contact_active", "contact_R"):
            a = getattr(batch, name)
            setattr(sub, name, None if a is None else np.ascontiguousarray(a[lo:hi]))
    return sub


def gather_solutions(local, B, dist, device=None, dst=None):
    """The single end-of-batch collective.  ``local``: dict of torch tensors for this rank's shard
    (X [b,N+1,9], U [b,N,nu], ints [3,b]).  Returns the full-batch tensors on every rank
    (all_gather) or on ``dst`` only (gather).  Shards may be ragged: they are padded to the
    largest shard for the collective and trimmed afterwards."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    sizes = [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]
    bmax = max(sizes)
    out = {}
    for key, t in local.items():
        lead = t.shape[-1] if key == "ints" else t.shape[0]
        if key == "ints":
            t = t.transpose(0, 1).contiguous()           # [b, 3]
        if lead == bmax:      # equal shards (the usual case): the shard itself is the send buffer
            pad = t.contiguous()
        else:
            pad = torch.zeros((bmax,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            pad[:lead] = t
        bufs = [torch.empty_like(pad) for _ in range(world)]
        if dst is None:
            dist.all_gather(bufs, pad)
        else:
            dist.gather(pad, bufs if rank == dst else None, dst=dst)
            if rank != dst:
                out[key] = None
                continue
        full = torch.cat([bufs[r][:sizes[r]] for r in range(world)], dim=0)
        out[key] = full.transpose(0, 1).contiguous() if key == "ints" else full
    return out
