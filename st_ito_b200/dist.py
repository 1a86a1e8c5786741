"""Population sharding for multi-GPU runs (one process per GPU, torch.distributed).

Candidates are independent given (input, chain, encoder, target), so rank r evaluates the contiguous
slice [r*ceil(P/G), ...) of the population and the only data-path collective is one all-gather of
the fp32 fitness values per generation (NCCL over NVLink on GPUs; gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np
import torch


def world():
    """(rank, world_size) of the default process group, (0, 1) when not initialised."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def backend() -> str:
    """Backend name of the default process group ("nccl", "gloo"), "" when not initialised."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized():
        return str(dist.get_backend())
    return ""


def shard_bounds(P: int, world_size: int, rank: int):
    """Contiguous slice [lo, hi) of a population of P owned by `rank`; chunk = ceil(P / world)."""
    chunk = -(-P // world_size)
    lo = min(rank * chunk, P)
    hi = min(lo + chunk, P)
    return lo, hi, chunk


def all_gather_rows(local: torch.Tensor, P: int, device=None) -> torch.Tensor:
    """Gather per-rank row blocks [n_r, ...] (n_r from shard_bounds) into [P, ...] on every rank.

    Rows are padded to the common chunk size so a single all_gather_into_tensor suffices.
    """
    import torch.distributed as dist

    rank, ws = world()
    if ws == 1:
        return local
    _, _, chunk = shard_bounds(P, ws, rank)
    backend = dist.get_backend()
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    buf = torch.zeros((chunk,) + tuple(local.shape[1:]), dtype=local.dtype, device=device)
    if local.shape[0] > 0:
        buf[: local.shape[0]].copy_(local, non_blocking=True)
    out = torch.empty((ws * chunk,) + tuple(local.shape[1:]), dtype=local.dtype, device=device)
    dist.all_gather_into_tensor(out, buf)
    return out[:P]


def broadcast_array(a: np.ndarray, src: int = 0) -> np.ndarray:
    """Broadcast a float64 array of identical shape on every rank from `src`."""
    import torch.distributed as dist

    rank, ws = world()
    if ws == 1:
        return a
    backend = dist.get_backend()
    device = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    t = torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(device)
    dist.broadcast(t, src=src)
    return t.cpu().numpy()
