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


def _gather_indexed(done: dict, n_total: int, seal_words: int, device: torch.device) -> list[np.ndarray] | None:
    """Gather {segment index: seal} of irregular shares (after a re-assignment): rows of [index, seal words], -1 = padding."""
    world, rank = dist.get_world_size(), dist.get_rank()
    cnt = torch.tensor([len(done)], dtype=torch.int64, device=device)
    dist.all_reduce(cnt, op=dist.ReduceOp.MAX)
    rows = int(cnt.item())
    stage = torch.full((rows, 1 + seal_words), -1, dtype=torch.int32)
    st = stage.numpy()
    for k, (i, s) in enumerate(sorted(done.items())):
        st[k, 0] = i
        st[k, 1:] = np.asarray(s, dtype=np.uint32).view(np.int32)
    allg = torch.empty((world * rows, 1 + seal_words), dtype=torch.int32, device=device)
    dist.all_gather_into_tensor(allg, stage.to(device))
    if rank != 0:
        return None
    host = allg.cpu().numpy()
    seals: list = [None] * n_total
    for row in host:
        if row[0] >= 0:
            seals[int(row[0])] = row[1:].view(np.uint32).copy()
    missing = [i for i, s in enumerate(seals) if s is None]
    if missing:
        raise RuntimeError(f"seals of segments {missing} are missing after the gather")
    return seals


def prove_sharded(n_total: int, prove_fn, seal_words: int, device: torch.device, max_attempts: int = 2):
    """The N-GPU session: segment i -> rank i mod world, every rank proves its share with `prove_fn(indices) -> [seal, ...]`
    (b2z_prove_session on that rank's GPU), rank 0 gets the seals back in segment order.

    Failure handling (upstream has none -- a failed segment fails `prove_session`; a node of 8 GPUs should not lose a block's
    session to one bad device): b2z_prove_session already retries a device error once on the same GPU; if a rank's
    `prove_fn` still raises, that rank's whole share is re-assigned round-robin to the ranks that succeeded and proved
    there (one extra collective, a MAX all-reduce of the per-segment failure flags, per attempt).  When no attempt is left
    or no healthy rank exists, every rank raises the same error.  Returns (seals in order on rank 0 / None elsewhere,
    number of segments that were re-assigned)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    if world == 1:
        return [np.asarray(s, dtype=np.uint32) for s in prove_fn(list(range(n_total)))], 0
    todo = partition(n_total, world, rank)
    done: dict[int, np.ndarray] = {}
    reassigned = 0
    bad_ranks: set[int] = set()
    for attempt in range(max_attempts):
        failed_here, why = [], ""
        if todo:
            try:
                seals = prove_fn(todo)
                if len(seals) != len(todo):
                    raise RuntimeError(f"prove_fn returned {len(seals)} seals for {len(todo)} segments")
                done.update(zip(todo, seals))
            except Exception as e:  # noqa: BLE001 -- any failure of this rank's prover re-assigns its share
                failed_here, why = todo, f"{type(e).__name__}: {e}"
        flags = torch.zeros(n_total + world, dtype=torch.int32, device=device)
        for i in failed_here:
            flags[i] = 1
        if failed_here:
            flags[n_total + rank] = 1
        dist.all_reduce(flags, op=dist.ReduceOp.MAX)
        flags = flags.cpu().numpy()
        failed = [int(i) for i in np.nonzero(flags[:n_total])[0]]
        if not failed:
            break
        bad_ranks |= {r for r in range(world) if flags[n_total + r]}
        healthy = [r for r in range(world) if r not in bad_ranks]
        if attempt + 1 == max_attempts or not healthy:
            raise RuntimeError(f"segments {failed} failed on rank(s) {sorted(bad_ranks)} and could not be re-assigned"
                               + (f" (this rank: {why})" if why else ""))
        todo = [seg for k, seg in enumerate(failed) if healthy[k % len(healthy)] == rank]
        reassigned += len(failed)
    if reassigned == 0:
        return all_gather_seals([done[i] for i in partition(n_total, world, rank)], n_total, seal_words, device), 0
    return _gather_indexed(done, n_total, seal_words, device), reassigned
