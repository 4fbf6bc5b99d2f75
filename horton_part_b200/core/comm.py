"""Communicators of the sharded (multi-GPU) runs.

``comm=`` of the partitioning classes accepts either a ``torch.distributed`` process group (NCCL backend
on the GPUs, gloo in the CPU tests of the sharding logic) or an :class:`HpComm` -- an NCCL communicator
created and driven entirely through the C ABI (``hp_comm_*`` in include/hp_b200.h), for callers that use
``libhp_b200.so`` without ``torch.distributed``.  The helpers below hide the difference; every collective
of the path is a sum / max / min all-reduce of a small FP64 vector."""

from __future__ import annotations

import numpy as np

from .. import _lib

__all__ = ["HpComm", "all_reduce", "comm_rank", "comm_world"]


class HpComm:
    """NCCL communicator behind the C ABI: one process per GPU, rank 0 calls :meth:`unique_id` and ships the
    128 bytes to the other ranks by whatever host channel the launcher has (file, socket, MPI, a
    ``torch.distributed`` store), then every rank constructs ``HpComm(world, rank, id, device)``."""

    def __init__(self, world, rank, unique_id, device=None):
        import torch

        if len(unique_id) != 128:
            raise ValueError("an NCCL unique id has 128 bytes")
        self.world, self.rank = int(world), int(rank)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        out = np.zeros(1, dtype=np.uint64)
        with torch.cuda.device(self.device):
            _lib.call("hp_comm_init", self.world, self.rank, np.frombuffer(bytes(unique_id), dtype=np.uint8).copy(), out)
        self._handle = int(out[0])

    @staticmethod
    def unique_id() -> bytes:
        buf = np.zeros(128, dtype=np.uint8)
        _lib.call("hp_comm_unique_id", buf)
        return buf.tobytes()

    def all_reduce(self, tensor, op="sum"):
        """In-place all-reduce of a contiguous FP64 device tensor on the current stream."""
        import torch

        if tensor.dtype != torch.float64 or not tensor.is_contiguous():
            raise TypeError("HpComm.all_reduce needs a contiguous float64 tensor")
        stream = torch.cuda.current_stream(tensor.device).cuda_stream
        if op == "min":  # min(x) = -max(-x)
            tensor.neg_()
            _lib.call("hp_comm_allreduce", self._handle, tensor, tensor.numel(), 1, stream)
            tensor.neg_()
        else:
            _lib.call("hp_comm_allreduce", self._handle, tensor, tensor.numel(), int(op == "max"), stream)
        return tensor

    def close(self):
        if getattr(self, "_handle", 0):
            _lib.call("hp_comm_destroy", self._handle)
            self._handle = 0

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass


def comm_rank(comm):
    if isinstance(comm, HpComm):
        return comm.rank
    import torch.distributed as dist

    return dist.get_rank(comm)


def comm_world(comm):
    if isinstance(comm, HpComm):
        return comm.world
    import torch.distributed as dist

    return dist.get_world_size(comm)


def all_reduce(comm, tensor, op="sum"):
    """In-place all-reduce over ``comm`` (``op``: "sum", "max" or "min")."""
    if isinstance(comm, HpComm):
        import torch

        if tensor.dtype == torch.float64 and tensor.is_contiguous():
            return comm.all_reduce(tensor, op)
        tmp = tensor.to(torch.float64).contiguous()
        comm.all_reduce(tmp, op)
        tensor.copy_(tmp)
        return tensor
    import torch.distributed as dist

    ops = {"sum": dist.ReduceOp.SUM, "max": dist.ReduceOp.MAX, "min": dist.ReduceOp.MIN}
    dist.all_reduce(tensor, op=ops[op], group=comm)
    return tensor
