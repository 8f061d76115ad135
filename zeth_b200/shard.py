"""One-process-per-GPU segment sharding.

Segments of a session are independent (upstream proves them in a sequential loop and concatenates the
SegmentReceipts in index order), so the path shards with NO data-path collective: segment i goes to rank
i mod world.  torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests) is used only to
broadcast the segment table from rank 0 and to gather the seals back to rank 0.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def partition(n_segments: int, world: int, rank: int) -> list[int]:
    """Round-robin: the segment indices rank `rank` proves."""
    return list(range(rank, n_segments, world))


def broadcast_segment_table(table: np.ndarray | None, device: torch.device) -> np.ndarray:
    """Rank 0 holds an int64 table [n_segments, k] (index, seed, po2, ...); every rank receives a copy."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return np.asarray(table, dtype=np.int64)
    shape = torch.zeros(2, dtype=torch.int64, device=device)
    if dist.get_rank() == 0:
        t = torch.as_tensor(np.asarray(table, dtype=np.int64)).to(device)
        shape[0], shape[1] = t.shape
    dist.broadcast(shape, 0)
    if dist.get_rank() != 0:
        t = torch.zeros(int(shape[0]), int(shape[1]), dtype=torch.int64, device=device)
    dist.broadcast(t, 0)
    return t.cpu().numpy()


def gather_seals(local: dict[int, np.ndarray], n_segments: int, seal_words: int, device: torch.device) -> list[np.ndarray] | None:
    """Gather {segment index: seal} from every rank to rank 0, returned in segment order (None on other ranks).

    All seals of one (circuit, po2) have the same length, so a fixed-shape all_gather of
    [max_local, 1 + seal_words] int32 rows (first word = segment index, -1 = padding) is enough."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    if world == 1:
        return [local[i] for i in range(n_segments)]
    max_local = (n_segments + world - 1) // world
    buf = torch.full((max_local, 1 + seal_words), -1, dtype=torch.int32)
    for k, (idx, seal) in enumerate(sorted(local.items())):
        buf[k, 0] = idx
        buf[k, 1:] = torch.from_numpy(seal.view(np.int32))
    buf = buf.to(device)
    out = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    dist.gather(buf, out, dst=0)
    if rank != 0:
        return None
    seals: list = [None] * n_segments
    for t in out:
        t = t.cpu().numpy()
        for row in t:
            if row[0] >= 0:
                seals[int(row[0])] = row[1:].view(np.uint32).copy()
    assert all(s is not None for s in seals), "a segment's seal is missing"
    return seals



def all_gather_seals(local_seals: list, n_total: int, seal_words: int, device: torch.device) -> list[np.ndarray] | None:
    """The block job's seal gather: of `n_total` segments rank r holds the seals of segments r, r + world, r + 2 world, ... in that
    order; ONE all_gather_into_tensor of exact seal words from a pinned staging buffer brings every seal to every rank, and
    rank 0 returns them in segment order (None on the other ranks).  Shorter shares are padded with zero rows."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    if world == 1:
        return [np.asarray(s, dtype=np.uint32) for s in local_seals]
    per_rank = (n_total + world - 1) // world
    assert len(local_seals) == len(range(rank, n_total, world))
    stage = torch.zeros((per_rank, seal_words), dtype=torch.int32)
    if device.type == "cuda":
        stage = stage.pin_memory()
    st = stage.numpy().view(np.uint32)
    for i, s in enumerate(local_seals):
        st[i] = s
    mine = stage.to(device, non_blocking=True)
    allg = torch.empty((world * per_rank, seal_words), dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(allg, mine)
    if rank != 0:
        return None
    host = allg.cpu().numpy().view(np.uint32)
    return [host[(i % world) * per_rank + i // world] for i in range(n_total)]
