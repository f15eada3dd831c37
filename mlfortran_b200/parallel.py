"""Sharding of the EM-GMM path over ranks (one process per GPU, torch.distributed for the plumbing).

The reference has no distributed layer (OpenMP over components only, src/unsupervised/mlf_emgmm.f90:179,199).
Every E/M quantity is a sum over points, so points are partitioned into contiguous shards and the engine
all-reduces ONE small fp64 buffer per iteration on the single-sweep path -- pass-A statistics, thresholded scatter and
the side-list flags together, 144 KB at d=16, K=64 -- plus a second, tiny one only when pairs were listed (the two-pass
shapes, d > 16, reduce pass-A statistics and scatter separately), and the global moments once per data set.

Two ways to supply the collective, both leaving the decision of WHAT to reduce to the engine:
  * attach_nccl(em, ...)   native: the library dlopens libnccl.so.2, builds its own communicator from a 128-byte id
                           that torch.distributed only has to broadcast once, and every iteration's all-reduce is a
                           plain ncclAllReduce on the engine's stream -- no Python in the iteration loop;
  * make_allreduce(...)    a Python callback over torch.distributed (gloo for the CPU tests of the host logic, NCCL as
                           a fallback).
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


def attach_nccl(em, rank: int, world: int, group=None, device=None) -> None:
    """Native NCCL all-reduce for a one-process-per-GPU job: rank 0 draws an ncclUniqueId through the library
    (mlf_b200_nccl_unique_id), torch.distributed broadcasts those 128 bytes once, and every rank joins with
    mlf_b200_nccl_init.  A C or Fortran caller does the same with MPI_Bcast."""
    import torch
    import torch.distributed as dist

    from . import _lib

    lib = _lib.load()
    buf = (C.c_ubyte * 128)()
    if rank == 0 and lib.mlf_b200_nccl_unique_id(buf) != 0:
        raise RuntimeError("mlf_b200_nccl_unique_id failed (libnccl.so.2 not loadable?)")
    t = torch.tensor(list(buf), dtype=torch.uint8, device=device if device is not None else "cpu")
    dist.broadcast(t, src=0, group=group)
    ident = (C.c_ubyte * 128)(*t.cpu().tolist())
    rc = lib.mlf_b200_nccl_init(em._h, ident, rank, world)
    if rc != 0:
        raise RuntimeError(f"mlf_b200_nccl_init failed with {rc}")
