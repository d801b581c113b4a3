"""Pair-sharded multi-GPU execution.

Every (source, target) registration is independent, so the batched workloads shard by pair: rank r
of R takes a contiguous block of ceil(P/R) pairs, aligns it on its own GPU with no inter-GPU traffic,
and the fixed 96-byte result records are all-gathered at the end (NCCL over NVLink on GPUs; the same
code runs over gloo on CPU for tests). No other collective exists on this path.
"""
from __future__ import annotations

import numpy as np


def shard_range(n_pairs: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block of pairs [begin, end) owned by ``rank``; blocks of ceil(P/R), the tail ranks may be empty."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    per = -(-n_pairs // world) if n_pairs > 0 else 0
    b = min(n_pairs, rank * per)
    return b, min(n_pairs, b + per)


def shard_segments(n_scans: int, rank: int, world: int) -> tuple[int, int]:
    """Sequential odometry over ``n_scans`` scans (n_scans - 1 pairs): rank r owns pairs [b, e) and therefore
    needs scans [b, e]; guesses chain only inside a segment so results do not depend on R."""
    b, e = shard_range(max(n_scans - 1, 0), rank, world)
    return b, e


def all_gather_results(local: np.ndarray, n_pairs: int, rank: int, world: int, device=None) -> np.ndarray:
    """All-gather the per-rank structured result arrays (RESULT_DTYPE, 96 B records) into pair order.

    ``local`` must hold exactly the records of shard_range(n_pairs, rank, world). Works on any
    initialised torch.distributed backend; with NCCL pass the rank's CUDA device.
    """
    from .fast_apdgicp import RESULT_DTYPE
    if world == 1:
        return local
    import torch
    import torch.distributed as dist
    per = -(-n_pairs // world)
    b, e = shard_range(n_pairs, rank, world)
    if local.shape[0] != e - b:
        raise ValueError("local shard has the wrong number of records")
    buf = np.zeros(per, dtype=RESULT_DTYPE)
    buf[: e - b] = local
    t = torch.from_numpy(buf.view(np.uint8).reshape(-1).copy())
    if device is not None:
        t = t.to(device)
    out = torch.empty(world * t.numel(), dtype=torch.uint8, device=t.device)
    dist.all_gather_into_tensor(out, t)
    allrec = np.frombuffer(out.cpu().numpy().tobytes(), dtype=RESULT_DTYPE).reshape(world, per)
    parts = []
    for r in range(world):
        rb, re_ = shard_range(n_pairs, r, world)
        parts.append(allrec[r, : re_ - rb])
    return np.concatenate(parts) if parts else np.zeros(0, dtype=RESULT_DTYPE)


def align_pairs_sharded(align_fn, n_pairs: int, rank: int, world: int, device=None) -> np.ndarray:
    """Run ``align_fn(begin, end) -> RESULT_DTYPE[end-begin]`` on this rank's block and gather everything."""
    b, e = shard_range(n_pairs, rank, world)
    local = align_fn(b, e)
    return all_gather_results(np.asarray(local), n_pairs, rank, world, device)
