"""One-process-per-GPU plumbing over torch.distributed (NCCL on the GPU box, gloo in CPU tests).

The encode path shards signals (columns of X) over ranks with NO collective
(SURVEY.md §8e).  The learners need:
  * ODL: one all-reduce(sum) of the sufficient statistics A (K,K) and B (n,K) per minibatch —
    exact because the block update is Jacobi (lyssa/dict_learning/online_dict_learn.py:91-94);
  * approximate K-SVD: a scalar all-reduce of ||X - DZ||^2 per iteration here, and the
    per-atom (n+2)-float reduction INSIDE the sweep kernel over peer-mapped buffers
    (lys_comm_* in include/lyssa_b200.h), whose 64-byte handles are exchanged with
    ``allgather_bytes``."""
from __future__ import annotations

import os

import torch
import torch.distributed as dist

from .utils import shard_bounds


class DistContext(object):
    def __init__(self, rank=0, world=1, group=None):
        self.rank, self.world, self.group = rank, world, group

    @classmethod
    def from_env_or_group(cls, group=None):
        if dist.is_available() and dist.is_initialized():
            return cls(dist.get_rank(group), dist.get_world_size(group), group)
        return cls(0, 1, None)

    @classmethod
    def init_from_env(cls, backend=None):
        """torchrun-style init (RANK / WORLD_SIZE / LOCAL_RANK / MASTER_*)."""
        world = int(os.environ.get("WORLD_SIZE", "1"))
        if world == 1:
            return cls(0, 1, None)
        if not dist.is_initialized():
            local = int(os.environ.get("LOCAL_RANK", "0"))
            if backend is None:
                backend = "nccl" if torch.cuda.is_available() else "gloo"
            if backend == "nccl":
                torch.cuda.set_device(local)
                dist.init_process_group(backend, device_id=torch.device("cuda", local))
            else:
                dist.init_process_group(backend)
        return cls(dist.get_rank(), dist.get_world_size(), None)

    def shard(self, n_columns):
        return shard_bounds(n_columns, self.world, self.rank)

    def allreduce_sum_(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def allreduce_max_(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX, group=self.group)
        return t

    def barrier(self):
        if self.world > 1:
            dist.barrier(group=self.group)

    def broadcast_(self, t, src=0):
        if self.world > 1:
            dist.broadcast(t, src=src, group=self.group)
        return t

    def allgather_object(self, obj):
        """every rank's (picklable) object, in rank order"""
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        dist.all_gather_object(out, obj, group=self.group)
        return out

    def broadcast_object(self, obj, src=0):
        if self.world == 1:
            return obj
        box = [obj if self.rank == src else None]
        dist.broadcast_object_list(box, src=src, group=self.group)
        return box[0]

    def allgather_bytes(self, payload: bytes):
        if self.world == 1:
            return [payload]
        out = [None] * self.world
        dist.all_gather_object(out, payload, group=self.group)
        return out


class PeerExchange(object):
    """Peer-mapped exchange buffers for the in-kernel per-atom all-reduce of the K-SVD sweep
    (lys_comm_* in include/lyssa_b200.h): every rank creates its buffer, the 64-byte CUDA IPC
    handles are all-gathered through torch.distributed, and each rank maps its peers' buffers.
    ``handle`` is the opaque pointer to pass as ``comm`` to engine.approx_ksvd_sweep."""

    def __init__(self, ctx: DistContext):
        import ctypes
        from . import _native as nat
        self.ctx = ctx
        self.handle = None
        if ctx.world == 1:
            return
        lib = nat.load()
        out = ctypes.c_void_p()
        nat.check(lib.lys_comm_create(ctx.rank, ctx.world, ctypes.byref(out)))
        self.handle = out.value
        buf = (ctypes.c_ubyte * nat.COMM_HANDLE_BYTES)()
        nat.check(lib.lys_comm_export(ctypes.c_void_p(self.handle), buf))
        handles = ctx.allgather_bytes(bytes(buf))
        blob = b"".join(handles)
        arr = (ctypes.c_ubyte * len(blob)).from_buffer_copy(blob)
        nat.check(lib.lys_comm_connect(ctypes.c_void_p(self.handle), arr))
        ctx.barrier()

    def close(self):
        if self.handle is not None:
            from . import _native as nat
            torch.cuda.synchronize()
            self.ctx.barrier()
            nat.load().lys_comm_destroy(__import__("ctypes").c_void_p(self.handle))
            self.handle = None
