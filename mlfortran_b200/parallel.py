"""Sharding of the EM-GMM path over ranks (one process per GPU, torch.distributed for the plumbing).

The reference has no distributed layer (OpenMP over components only, src/unsupervised/mlf_emgmm.f90:179,199).
Every E/M quantity is a sum over points, so points are partitioned into contiguous shards and the engine
all-reduces two small fp64 buffers per iteration (pass-A statistics K*(2+d), pass-B scatter K*d*d) plus the
global moments once.  This module only supplies the partition arithmetic and the all-reduce callback; the
engine decides what to reduce.
"""
from __future__ import annotations

import ctypes as C

import numpy as np


def shard_bounds(n_total: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block partition [lo, hi) of n_total points; the first n_total % world ranks get one extra."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, rem = divmod(int(n_total), world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


class _DevBuf:
    """Minimal __cuda_array_interface__ carrier so torch can alias a raw device pointer without copying."""

    def __init__(self, ptr: int, count: int):
        self.__cuda_array_interface__ = {"shape": (count,), "typestr": "<f8", "data": (ptr, False), "version": 3, "strides": None}


def make_allreduce(group=None, device=None):
    """Returns fn(buf_ptr, count, stream) -> 0 performing an in-place SUM all-reduce of `count` doubles.

    device=None / "cpu": buf_ptr is host memory (gloo; used by the CPU tests of the host logic).
    device=cuda:i     : buf_ptr is device memory and the collective is ordered on `stream` (NCCL)."""
    import torch
    import torch.distributed as dist

    on_cuda = device is not None and str(device).startswith("cuda")

    def fn(ptr: int, count: int, stream: int) -> int:
        if count <= 0:
            return 0
        if on_cuda:
            t = torch.as_tensor(_DevBuf(ptr, count), device=device)
            ext = torch.cuda.ExternalStream(stream, device=device) if stream else torch.cuda.current_stream(device)
            with torch.cuda.stream(ext):
                dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        else:
            arr = np.ctypeslib.as_array(C.cast(ptr, C.POINTER(C.c_double)), shape=(count,))
            t = torch.from_numpy(arr)
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
        return 0

    return fn
