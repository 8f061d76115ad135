#!/usr/bin/env python3
"""Sweep the launch configuration of the mixed INT / FP64 hash_rows kernel (kernels_hash.cu) on one B200.

  python tools/sweep_hash.py [--po2 22] [--cols 64]
Each line: configuration, median ms, Poseidon2 permutations / s, SM clock sampled while the kernel loops."""
import argparse
import json
import os
import subprocess
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zeth_b200 import hal as H  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--po2", type=int, default=22)
ap.add_argument("--cols", type=int, default=64)
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
dev = torch.device("cuda", 0)
hal = H.Hal(0)
stream = torch.cuda.Stream(device=dev)
hal.set_stream(stream.cuda_stream)
P = 15 * (1 << 27) + 1
rows = 1 << a.po2
g = torch.Generator(device=dev)
g.manual_seed(5)
with torch.cuda.stream(stream):
    m = torch.randint(0, P, (a.cols * rows,), dtype=torch.int32, device=dev, generator=g)
    dig = torch.empty(8 * rows, dtype=torch.int32, device=dev)
    ref = torch.empty(8 * rows, dtype=torch.int32, device=dev)
torch.cuda.synchronize()
perms = rows * ((a.cols + 15) // 16)


def clocks():
    try:
        o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.active", "--format=csv,noheader,nounits", "-i", "0"],
                           capture_output=True, text=True, timeout=10).stdout.strip()
        return o
    except Exception as e:  # noqa: BLE001
        return str(e)


def run(cfg):
    for k, v in cfg.items():
        os.environ[k] = str(v)
    fn = lambda: hal.hash_rows(dig.data_ptr(), m.data_ptr(), rows, a.cols)  # noqa: E731
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(a.reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    # clocks / power under sustained load: queue ~1.5 s of launches and sample in the middle
    n_loop = max(4, int(1500 / np.median(ts)))
    for _ in range(n_loop):
        fn()
    import time
    time.sleep(0.7)
    c = clocks()
    torch.cuda.synchronize()
    ms = float(np.median(ts))
    return {"cfg": cfg, "ms": round(ms, 3), "Gperm_s": round(perms / ms / 1e6, 3), "clocks_power_reasons": c}


base = run({"B2Z_HASH_MIXED": 0})
ref.copy_(dig)
print(json.dumps(base), flush=True)
ap_cfgs = os.environ.get("SWEEP_CFGS", "6:2:1,6:3:2,7:2:1,7:3:2,7:7:4,7:7:5,8:2:1,8:4:3,8:8:5,8:8:6")
for c in ap_cfgs.split(","):
    ctas, period, n_int = (int(v) for v in c.split(":"))
    dig.zero_()
    r = run({"B2Z_HASH_MIXED": 1, "B2Z_HASH_CTAS": ctas, "B2Z_HASH_ROLE_PERIOD": period, "B2Z_HASH_ROLE_INT": n_int})
    torch.cuda.synchronize()
    r["same_digests"] = bool(torch.equal(dig, ref))
    print(json.dumps(r), flush=True)
hal.close()
