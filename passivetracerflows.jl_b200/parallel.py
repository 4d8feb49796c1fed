"""Multi-GPU plumbing: one process per GPU (torchrun / torch.distributed), the library's own NCCL communicator for
the only collective on the path (the all-to-all transpose inside slab-decomposed transforms).

torch.distributed is used ONLY to agree on the ncclUniqueId (broadcast of 128 bytes) and, in bench.py, for barriers
and the max-over-ranks reduction.  Works with the "nccl" backend on GPUs and the "gloo" backend in CPU tests.
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _capi
from .tracer_advection_diffusion import B200


def slab_extents(n_global: int, nranks: int, rank: int):
    """(count, offset) of the contiguous slab a rank owns along one axis (requires divisibility, as the library does)."""
    if n_global % nranks:
        raise ValueError(f"extent {n_global} is not divisible by {nranks} ranks")
    cnt = n_global // nranks
    return cnt, cnt * rank


def init_b200(decomposition: str = "slab", device: int | None = None, engine: str = "auto") -> B200:
    """Build the ``B200`` device descriptor of this rank.  Must be called by every rank (collective)."""
    import torch.distributed as dist
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    if device is None:
        device = int(os.environ.get("LOCAL_RANK", rank))
    nccl_id = None
    if world > 1 and decomposition == "slab":
        buf = np.zeros(128, dtype=np.uint8)
        if rank == 0:
            _capi.check(_capi.load().ptf_nccl_unique_id(buf.ctypes.data_as(C.POINTER(C.c_uint8))))
        obj = [buf.tobytes()]
        dist.broadcast_object_list(obj, src=0)
        nccl_id = obj[0]
    return B200(device=device, engine=engine, rank=rank, nranks=world, nccl_id=nccl_id,
                decomposition=decomposition if world > 1 else "none")
