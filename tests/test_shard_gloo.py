"""N > 1 host logic on CPU: world_size-2 gloo run of the segment-table broadcast, the round-robin partition and the
seal gather.  The seals here come from the oracle (no GPU in this test)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from zeth_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_segments, seal_words, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    dev = torch.device("cpu")
    table = np.array([[i, 1000 + i, 10] for i in range(n_segments)], dtype=np.int64) if rank == 0 else None
    table = shard.broadcast_segment_table(table, dev)
    mine = shard.partition(n_segments, world, rank)
    local = {i: (np.arange(seal_words, dtype=np.uint32) * np.uint32(table[i][1])) for i in mine}
    seals = shard.gather_seals(local, n_segments, seal_words, dev)
    # the block job's gather (bench.py): equal shares, rank r holds segments r, r + world, ... in that order
    mine2 = [i for i in range(n_segments) if i % world == rank]  # 5 segments over 2 ranks: ragged shares
    seals2 = shard.all_gather_seals([np.arange(seal_words, dtype=np.uint32) * np.uint32(7 + i) for i in mine2], n_segments, seal_words, dev)
    if rank == 0:
        ok = all(np.array_equal(seals[i], np.arange(seal_words, dtype=np.uint32) * np.uint32(1000 + i)) for i in range(n_segments))
        ok = ok and len(seals2) == n_segments and all(
            np.array_equal(seals2[i], np.arange(seal_words, dtype=np.uint32) * np.uint32(7 + i)) for i in range(n_segments))
        q.put((ok, [int(t[1]) for t in table]))
    else:
        assert seals2 is None
    dist.destroy_process_group()


def test_partition_covers_every_segment_once():
    for world in (1, 2, 4, 8):
        for n in (1, 7, 500):
            parts = [shard.partition(n, world, r) for r in range(world)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_broadcast_and_gather_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 5, 64, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    ok, seeds = q.get(timeout=5)
    assert ok and seeds == [1000, 1001, 1002, 1003, 1004]


def _retry_worker(rank, world, port, n_segments, seal_words, fail_rank, fail_always, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    calls = []

    def prove_fn(indices):
        calls.append(list(indices))
        if rank == fail_rank or fail_always:
            raise RuntimeError("device lost")
        return [np.arange(seal_words, dtype=np.uint32) * np.uint32(3 + i) for i in indices]

    try:
        seals, moved = shard.prove_sharded(n_segments, prove_fn, seal_words, torch.device("cpu"))
        if rank == 0:
            ok = all(np.array_equal(seals[i], np.arange(seal_words, dtype=np.uint32) * np.uint32(3 + i)) for i in range(n_segments))
            q.put(("ok", ok, moved, calls))
        else:
            assert seals is None
            q.put(("other", calls))
    except RuntimeError as e:
        q.put(("raised", rank, str(e)))
    dist.destroy_process_group()


def _run(world, args):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_retry_worker, args=(r, world, port) + args + (q,)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return out


def test_failed_ranks_share_is_reproved_on_the_healthy_rank():
    """rank 1's prover raises: its segments (1, 3) are re-assigned to rank 0 and the session still completes in order"""
    out = _run(2, (5, 32, 1, False))
    ok = [o for o in out if o[0] == "ok"][0]
    assert ok[1] is True and ok[2] == 2
    assert ok[3] == [[0, 2, 4], [1, 3]]                      # rank 0: its own share, then rank 1's
    assert [o for o in out if o[0] == "other"][0][1] == [[1, 3]]   # rank 1 is not asked again


def test_no_failure_takes_the_exact_gather():
    out = _run(2, (5, 32, -1, False))
    ok = [o for o in out if o[0] == "ok"][0]
    assert ok[1] is True and ok[2] == 0 and ok[3] == [[0, 2, 4]]


def test_every_rank_failing_raises_everywhere():
    out = _run(2, (4, 16, -1, True))
    assert sorted(o[0] for o in out) == ["raised", "raised"]
    assert all("could not be re-assigned" in o[2] for o in out)
