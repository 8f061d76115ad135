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

    def prove(self, segments, journal=b""):
        self.proved.append([s["index"] for s in segments])
        return _Receipt(len(segments), 40)

    def set_control_root(self, po2, root):
        pass

    def launch_count(self):
        return 100 * len(self.proved)

    def last_retries(self):
        return 0

    def close(self):
        pass


def test_block_workload_glue_runs_end_to_end_with_a_mocked_gpu(monkeypatch, capsys):
    from zeth_b200 import hal as H
    cpu = torch.device("cpu")
    real_device = torch.device
    monkeypatch.setattr(torch.cuda, "is_available", lambda: True)
    monkeypatch.setattr(torch.cuda, "set_device", lambda d: None)
    monkeypatch.setattr(torch.cuda, "synchronize", lambda *a: None)
    monkeypatch.setattr(torch.cuda, "empty_cache", lambda: None)
    monkeypatch.setattr(torch.cuda, "get_device_properties", lambda d: types.SimpleNamespace(total_memory=180 << 30))

    class _Event:
        def __init__(self, enable_timing=False):
            pass

        def record(self, *a):
            pass

        def elapsed_time(self, other):
            return 5.0

    monkeypatch.setattr(torch.cuda, "Event", _Event)
    monkeypatch.setattr(torch.cuda, "Stream", lambda device=None: types.SimpleNamespace(cuda_stream=0))
    monkeypatch.setattr(torch, "device", lambda *a, **k: cpu if (a and a[0] == "cuda") else real_device(*a, **k))
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    monkeypatch.setattr(H, "Session", _Session)
    monkeypatch.setattr(H.Receipt, "assemble", staticmethod(lambda circuit, seals, journal=b"": types.SimpleNamespace(close=lambda: None)))
    monkeypatch.setattr(bench, "kernel_rooflines", lambda *a, **k: ({"bound": "hbm", "frac": 0.1}, []))
    monkeypatch.setattr(bench, "cpu_reference_run", lambda steps, warmup, circuit="x": (1.0, 1, "mock", 1.0, 6))
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
