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


class PeerResults:
    """The solutions of all ranks in ONE buffer on the GPU of rank ``dst``, written by every rank's kernel at its
    tiles' write-back through NVLink peer memory (include/cmpc.h: cmpc_peer_alloc / cmpc_peer_open) -- the multi-GPU
    path without a gather: nothing runs on the SMs besides the solver.  ``slots`` result sets are allocated so that
    consecutive batches can alternate.  Collective: every rank constructs it (the handle is broadcast with
    ``dist.broadcast_object_list``)."""

    def __init__(self, B, N, nu, dist, dst=0, slots=2):
        import ctypes as C
        from . import _lib as L
        self.lib = L.load()
        self.B, self.N, self.nu, self.dst, self.slots = int(B), int(N), int(nu), dst, slots
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        nX, nU = self.B * (N + 1) * 9, self.B * N * nu
        self.off = {"X": 0, "U": nX * 8, "scp_iters": (nX + nU) * 8, "status": (nX + nU) * 8 + 4 * self.B,
                    "n_accepted": (nX + nU) * 8 + 8 * self.B}
        self.slot_bytes = ((nX + nU) * 8 + 12 * self.B + 255) // 256 * 256
        ptr, handle = C.c_void_p(), C.create_string_buffer(64)
        if self.rank == dst:
            L.check(self.lib.cmpc_peer_alloc(self.slot_bytes * slots, C.byref(ptr), handle), self.lib)
        box = [handle.raw if self.rank == dst else None]
        dist.broadcast_object_list(box, src=dst)
        if self.rank != dst:
            L.check(self.lib.cmpc_peer_open(box[0], C.byref(ptr)), self.lib)
        self.base = int(ptr.value)

    def pointers(self, slot, first):
        """Raw device pointers (valid in this process) of instance ``first`` onwards in result set ``slot``."""
        b = self.base + slot * self.slot_bytes
        N, nu = self.N, self.nu
        return {"X": b + self.off["X"] + first * (N + 1) * 9 * 8, "U": b + self.off["U"] + first * N * nu * 8,
                "scp_iters": b + self.off["scp_iters"] + 4 * first, "status": b + self.off["status"] + 4 * first,
                "n_accepted": b + self.off["n_accepted"] + 4 * first}

    def tensors(self, slot):
        """Zero-copy torch views of result set ``slot`` (on rank ``dst``, after every rank has synchronised)."""
        import torch
        if self.rank != self.dst:
            return None
        b = self.base + slot * self.slot_bytes

        class _View:
            def __init__(self, ptr, shape, typestr):
                self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 2}
        B, N, nu = self.B, self.N, self.nu
        dev = torch.device("cuda", torch.cuda.current_device())
        return {"X": torch.as_tensor(_View(b + self.off["X"], (B, N + 1, 9), "<f8"), device=dev),
                "U": torch.as_tensor(_View(b + self.off["U"], (B, N, nu), "<f8"), device=dev),
                "scp_iters": torch.as_tensor(_View(b + self.off["scp_iters"], (B,), "<i4"), device=dev),
                "status": torch.as_tensor(_View(b + self.off["status"], (B,), "<i4"), device=dev),
                "n_accepted": torch.as_tensor(_View(b + self.off["n_accepted"], (B,), "<i4"), device=dev)}

    def close(self):
        if self.base:
            (self.lib.cmpc_peer_free if self.rank == self.dst else self.lib.cmpc_peer_close)(self.base)
            self.base = 0
