"""Population sharding across GPUs: one process per GPU (torchrun), contiguous ranges of rollout units per rank,
and the per-generation exchange.

The reference's workers return only ``(noise_idx, returns)`` to the master (es.py:428-439) and the master forms
the gradient from the shared noise table; here every rank holds a replica of the table, theta and the optimizer
state, so per generation the ranks exchange
  1. all_gather of (returns, signreturns, lengths) -- a few KB -- so every rank computes IDENTICAL ranks,
  2. one all_reduce(sum) of the per-rank partial gradient (4*P bytes over NVLink / NVSwitch),
and then take the identical optimizer step locally (no theta broadcast).  GA: all_gather of fitness only
(genomes are seed chains every rank can rebuild).  Works with the ``nccl`` backend on CUDA tensors and ``gloo`` on
CPU tensors (the CPU test-suite drives the host logic through gloo at world_size 2).
"""
from __future__ import annotations

import os
from typing import List, Sequence, Tuple

import torch
import torch.distributed as dist


def dist_info() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_from_env(backend: str = None) -> Tuple[int, int, int]:
    """Join the torchrun process group if WORLD_SIZE > 1.  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        kw = {"device_id": torch.device("cuda", local)} if backend == "nccl" else {}
        dist.init_process_group(backend=backend, rank=rank, world_size=world, **kw)
    return rank, world, local


def shard_bounds(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous range [lo, hi) of unit indices owned by ``rank`` (both episodes of a pair stay on one GPU)."""
    return (n * rank) // world, (n * (rank + 1)) // world


def all_gather_rows(local: torch.Tensor, n_total: int) -> torch.Tensor:
    """Concatenate the per-rank row blocks (rank r holds rows shard_bounds(n_total, r, world)) in unit order."""
    rank, world = dist_info()
    if world == 1:
        assert local.shape[0] == n_total
        return local
    counts = [shard_bounds(n_total, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in counts)
    pad = torch.zeros((mx,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad.contiguous())
    return torch.cat([b[: hi - lo] for b, (lo, hi) in zip(bufs, counts)], dim=0)


def all_reduce_sum_(t: torch.Tensor) -> torch.Tensor:
    _, world = dist_info()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def broadcast_seed(seed) -> int:
    """Every rank must draw the same noise-index stream: rank 0's seed wins."""
    import numpy as np
    rank, world = dist_info()
    if seed is None:
        seed = int(np.random.RandomState().randint(2 ** 31 - 1))
    if world > 1:
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
        t = torch.tensor([seed], dtype=torch.int64, device=dev)
        dist.broadcast(t, src=0)
        seed = int(t.item())
    return seed


def broadcast_object(obj, src: int = 0):
    """Rank ``src``'s picklable object on every rank (identity at world size 1).  Used for small host-side results that
    must be IDENTICAL everywhere although every rank could recompute them (e.g. a behaviour characterisation from a
    possibly stochastic host environment)."""
    _, world = dist_info()
    if world == 1:
        return obj
    box = [obj]
    dist.broadcast_object_list(box, src=src)
    return box[0]


def all_gather_object(obj):
    """Every rank's picklable object, in rank order (small host-side records only: behaviour characterisations for the
    VINE export)."""
    _, world = dist_info()
    if world == 1:
        return [obj]
    out = [None] * world
    dist.all_gather_object(out, obj)
    return out


def barrier():
    _, world = dist_info()
    if world > 1:
        dist.barrier()
