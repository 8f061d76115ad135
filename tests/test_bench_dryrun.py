"""Host logic of bench.py's default (block) workload with the GPU layer mocked: the whole function runs top to bottom on the
CPU -- table broadcast, warm-up, resident-witness guard, clock-sampler marks, prove_sharded, receipt check, latency
statistics, the e2e leg, the JSON line -- so that an ordering or naming mistake in that glue cannot reach the GPU box.
Nothing here measures anything."""
import argparse
import json
import types

import numpy as np
import torch

import bench


class _Receipt:
    def __init__(self, n, words):
        self.n_segments, self.words = n, words

    def seal(self, i):
        return np.full(self.words, 7 + i, dtype=np.uint32)

    def close(self):
        pass


class _Hal:
    def __init__(self):
        self.launches = 0

    def witgen(self, circuit, po2, code, data, glob, noise, seed, zk):
        self.launches += 3

    def sync(self):
        pass

    def compute_control_root(self, po2, code, cols):
        return np.arange(8, dtype=np.uint32)

    def set_stream(self, s):
        pass

    def prove_segment_device(self, *a):
        return np.zeros(4, dtype=np.uint32)

    def last_phase_ms(self):
        return {"total": 1.0}


class _Circuit:
    def seal_words(self, po2):
        return 40

    def verify_receipt(self, receipt, image_id):
        return 0, 0


class _Lib:
    def b2z_session_last_segment_ms(self, h, buf, cap):
        for i in range(5):
            buf[i] = 10.0 + i
        return 5


class _Session:
    def __init__(self, device, blob, in_flight=2):
        self.hals, self.circuit, self.in_flight, self.L, self.h = [_Hal()], _Circuit(), in_flight, _Lib(), None
        self.proved = []

    fail_on_call = 0  # (tests) the N-th prove call of this process raises once, like a device that drops out

    def prove(self, segments, journal=b""):
        self.proved.append([s["index"] for s in segments])
        if len(self.proved) == _Session.fail_on_call:
            raise RuntimeError("segment 0: device lost")
        return _Receipt(len(segments), 40)

    def set_control_root(self, po2, root):
        pass

    def launch_count(self):
        return 100 * len(self.proved)

    def last_retries(self):
        return 0

    def close(self):
        pass


class _Event:
    def __init__(self, enable_timing=False):
        pass

    def record(self, *a):
        pass

    def elapsed_time(self, other):
        return 5.0


def _mock_gpu_layer(setattr_):
    """replace everything of torch.cuda / zeth_b200.hal that bench.run_block_workload touches (setattr_(obj, name, value))"""
    from zeth_b200 import hal as H
    cpu = torch.device("cpu")
    real_device = torch.device
    setattr_(torch.cuda, "is_available", lambda: True)
    setattr_(torch.cuda, "set_device", lambda d: None)
    setattr_(torch.cuda, "synchronize", lambda *a: None)
    setattr_(torch.cuda, "empty_cache", lambda: None)
    setattr_(torch.cuda, "get_device_properties", lambda d: types.SimpleNamespace(total_memory=180 << 30))
    setattr_(torch.cuda, "Event", _Event)
    setattr_(torch.cuda, "Stream", lambda device=None: types.SimpleNamespace(cuda_stream=0))
    setattr_(torch, "device", lambda *a, **k: cpu if (a and a[0] == "cuda") else real_device(*a, **k))
    setattr_(torch.Tensor, "pin_memory", lambda self: self)
    setattr_(H, "Session", _Session)
    setattr_(H.Receipt, "assemble", staticmethod(lambda circuit, seals, journal=b"": types.SimpleNamespace(close=lambda: None)))
    setattr_(bench, "kernel_rooflines", lambda *a, **k: ({"bound": "hbm", "frac": 0.1}, []))
    setattr_(bench, "cpu_reference_run", lambda steps, warmup, circuit="x": (1.0, 1, "mock", 1.0, 6))


def test_block_workload_glue_runs_end_to_end_with_a_mocked_gpu(monkeypatch, capsys):
    _mock_gpu_layer(monkeypatch.setattr)
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    monkeypatch.delenv("RANK", raising=False)
    args = argparse.Namespace(gpus=1, steps=3, warmup=1, workload="block", circuit="planning", po2=6, segments_per_step=2, in_flight=2,
                              impl="b2z", no_cpu_baseline=False)
    assert bench.run_block_workload(args) == 0
    line = json.loads(capsys.readouterr().out.strip().splitlines()[-1])
    assert line["metric"] == bench.METRIC and line["n_gpus"] == 1 and line["steps"] == 3 and line["warmup"] == 1
    cfg = line["config"]
    assert cfg["segments_total"] == 6 and cfg["distinct_witnesses_per_rank"] == 6 and cfg["segments_reassigned_across_ranks"] == 0
    assert cfg["receipt_verified_segments"] == 4 and cfg["segment_latency_ms"]["max"] == 14.0 and len(cfg["rank_prove_ms"]) == 1
    assert line["e2e"]["value"] > 0 and line["e2e"]["h2d_bytes_per_step"] > 0 and line["gpu_launches"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["roofline"]["bound"] == "hbm"


def _rank_main(rank, world, port, q, fail_rank=-1):
    import io
    import os
    import sys
    import torch.distributed as dist
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK="0", MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    _mock_gpu_layer(setattr)
    if rank == fail_rank:
        _Session.fail_on_call = 2  # call 1 = the warm-up, call 2 = this rank's timed share
    bench._init_dist = lambda local_rank: dist.init_process_group("gloo", rank=rank, world_size=world)
    args = argparse.Namespace(gpus=world, steps=3, warmup=1, workload="block", circuit="planning", po2=6, segments_per_step=3, in_flight=2,
                              impl="b2z", no_cpu_baseline=True)
    out, real = io.StringIO(), sys.stdout
    sys.stdout = out
    try:
        rc = bench.run_block_workload(args)
    finally:
        sys.stdout = real
    q.put((rank, rc, out.getvalue()))


def _run_two_ranks(fail_rank=-1):
    import socket
    import torch.multiprocessing as mp
    sock = socket.socket()
    sock.bind(("127.0.0.1", 0))
    port = sock.getsockname()[1]
    sock.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_rank_main, args=(r, 2, port, q, fail_rank)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict()
    for _ in range(2):
        rank, rc, text = q.get(timeout=180)
        got[rank] = (rc, text)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    return got


def test_a_rank_that_fails_its_timed_share_is_covered_by_the_other_rank():
    """rank 1's session call raises once: its 4 segments are regenerated and proved on rank 0 inside the timed region, the job
    still completes, and the line says how many segments moved"""
    got = _run_two_ranks(fail_rank=1)
    assert got[0][0] == 0 and got[1][0] == 0
    line = json.loads(got[0][1].strip().splitlines()[-1])
    assert line["config"]["segments_total"] == 9 and line["config"]["segments_reassigned_across_ranks"] == 4


def test_block_workload_glue_on_two_ranks_over_gloo():
    """the N > 1 branches of the same glue: table broadcast, warmed gathers, prove_sharded (ragged shares: 9 segments over 2
    ranks), per-rank time gather, the e2e gather; rank 0 alone prints the line"""
    got = _run_two_ranks()
    assert got[0][0] == 0 and got[1][0] == 0 and got[1][1].strip() == ""
    line = json.loads(got[0][1].strip().splitlines()[-1])
    assert line["n_gpus"] == 2 and line["config"]["segments_total"] == 9 and len(line["config"]["rank_prove_ms"]) == 2
    assert line["config"]["segments_reassigned_across_ranks"] == 0 and line["e2e"]["value"] > 0 and line["cpu_baseline"] is None
