#!/usr/bin/env python3
"""BASELINE.json config 5: BabyBear NTT and Poseidon2 hash_rows micro-benchmarks, achieved GB/s vs the measured HBM peak.

  python tools/bench_micro.py            # (po2, cols) in {(20,256), (22,256), (24,64)}, seed 0xB2000005
Prints one JSON object; every number is the median of `--reps` CUDA-event timings after 3 warm-ups, inputs >> L2."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zeth_b200 import hal as H  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--reps", type=int, default=10)
ap.add_argument("--shapes", default="20x256,22x256,24x64")
a = ap.parse_args()
peaks = {}
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    pass
hbm = float(peaks.get("hbm_gbs", 6650.0))
dev = torch.device("cuda", 0)
hal = H.Hal(0)
stream = torch.cuda.Stream(device=dev)
hal.set_stream(stream.cuda_stream)
P = 15 * (1 << 27) + 1
out = {"metric": "config-5 microbench", "hbm_peak_gbs": hbm, "peak_source": "measured" if "hbm_gbs" in peaks else "fallback", "rows": []}


def med(fn, reps):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        fn()
        e1.record(stream)
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return float(np.median(ts)), float(min(ts))


for shape in a.shapes.split(","):
    po2, cols = (int(x) for x in shape.split("x"))
    n = 1 << po2
    g = torch.Generator(device=dev)
    g.manual_seed(0xB2000005)
    with torch.cuda.stream(stream):
        x = torch.randint(0, P, (cols * n,), dtype=torch.int32, device=dev, generator=g)
        lde = torch.empty(cols * 4 * n, dtype=torch.int32, device=dev)
        dig = torch.empty(8 * 4 * n, dtype=torch.int32, device=dev)
    torch.cuda.synchronize()
    for name, fn, nbytes in (
        ("batch_expand_into_evaluate_ntt", lambda: hal.batch_expand_into_evaluate_ntt(lde.data_ptr(), x.data_ptr(), cols, po2, 2), 20 * cols * n),
        ("batch_interpolate_ntt", lambda: hal.batch_interpolate_ntt(x.data_ptr(), cols, po2), 8 * cols * n),
        ("hash_rows", lambda: hal.hash_rows(dig.data_ptr(), lde.data_ptr(), 4 * n, cols), 4 * cols * 4 * n + 32 * 4 * n),
    ):
        m, best = med(fn, a.reps if name != "hash_rows" else max(3, a.reps // 3))
        row = {"op": name, "po2": po2, "cols": cols, "algorithmic_bytes": nbytes, "ms_median": m, "ms_best": best,
               "achieved_gbs": nbytes / m / 1e6, "frac_hbm": nbytes / m / 1e6 / hbm}
        if name == "hash_rows":
            perms = 4 * n * ((cols + 15) // 16)
            row["poseidon2_perms_per_s"] = perms / (m * 1e-3)
            row["binding"] = "INT32 pipes (1356 modular products per permutation)"
        else:
            row["binding"] = "INT32 pipes (modular butterflies), then L2"
        out["rows"].append(row)
    del x, lde, dig
    torch.cuda.empty_cache()
print(json.dumps(out))
hal.close()
