"""Process-group plumbing for the row-sharded (one process per GPU) path.

torch.distributed is used only to launch/synchronise ranks and to hand the 128-byte
ncclUniqueId from rank 0 to the others; the data path (all-gather of the Krylov vector
inside GMRES) is NCCL called from libbemb200 itself.
"""
from __future__ import annotations

import os
from typing import Tuple

import numpy as np


def env_rank() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment (defaults: single process)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))


def init_process_group(backend: str):
    import torch.distributed as dist

    rank, _, world = env_rank()
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world


def broadcast_bytes(payload: bytes | None, nbytes: int, src: int = 0, device="cpu") -> bytes:
    """Broadcast a small byte string (the ncclUniqueId) from `src` to every rank."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        assert payload is not None
        return payload
    t = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    if dist.get_rank() == src:
        t.copy_(torch.frombuffer(bytearray(payload), dtype=torch.uint8))
    dist.broadcast(t, src=src)
    return bytes(t.cpu().numpy().tobytes())


def partition(n: int, nranks: int, rank: int) -> Tuple[int, int]:
    """Same split as bemb200_partition (csrc/api.cu): rank r owns [r*chunk, min(n,(r+1)*chunk))."""
    chunk = (n + nranks - 1) // nranks
    b = min(n, chunk * rank)
    e = min(n, b + chunk)
    return b, e


def gather_layout(n: int, nranks: int) -> Tuple[int, int]:
    """(chunk, padded length) of the all-gather buffer used for the Krylov vector (csrc/gmres.cu)."""
    chunk = (n + nranks - 1) // nranks
    return chunk, chunk * nranks


def allgather_rows(local: np.ndarray, n: int, backend_device="cpu") -> np.ndarray:
    """All-gather row slices (padded to `chunk`) into the full length-n vector, the way
    gmres.cu's matvec() does with ncclAllGather."""
    import torch
    import torch.distributed as dist

    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    chunk, npad = gather_layout(n, world)
    send = torch.zeros(chunk, dtype=torch.complex128)
    send[: local.shape[0]] = torch.from_numpy(local)
    out = torch.zeros(npad, dtype=torch.complex128)
    dist.all_gather_into_tensor(torch.view_as_real(out), torch.view_as_real(send))
    return out.numpy()[:n].copy()
