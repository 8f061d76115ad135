#!/usr/bin/env python3
"""Time eval_check (and, optionally, a whole segment proof) of a named circuit with CUDA events.
usage: python tools/prof_evalcheck.py [--circuit rv32im_scale|planning] [--po2 20] [--reps 3] [--prove]
Environment knobs (read at circuit-load time): B2Z_EVALCHECK_THREADS, B2Z_EVALCHECK_PARTS, B2Z_EVALCHECK_MINBLOCKS."""
import argparse
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zeth_b200 import hal as H, synth_circuit as S  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--circuit", default="rv32im_scale")
ap.add_argument("--po2", type=int, default=20)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--prove", action="store_true")
a = ap.parse_args()
sc = {"rv32im_scale": S.circuit_rv32im_scale, "planning": S.circuit_planning, "mux_small": S.circuit_mux_small}[a.circuit]()
dev = torch.device("cuda", 0)
hal = H.Hal(0)
stream = torch.cuda.Stream(device=dev)
hal.set_stream(stream.cuda_stream)
c = H.Circuit(hal, sc.blob)
po2, n = a.po2, 1 << a.po2
wa, wc, wd = sc.group_sizes
P = 15 * (1 << 27) + 1
with torch.cuda.stream(stream):
    groups = [torch.randint(0, P, (w * 4 * n,), dtype=torch.int32, device=dev) for w in (wa, wc, wd)]
    check = torch.empty(16 * n, dtype=torch.int32, device=dev)
    glob = torch.randint(0, P, (max(sc.globals_size, 1),), dtype=torch.int32, device=dev)
    mix = torch.randint(0, P, (max(sc.mix_size, 1),), dtype=torch.int32, device=dev)
pm = np.array([3, 5, 7, 11], dtype=np.uint32)


def run():
    hal.eval_check(c, check.data_ptr(), [g.data_ptr() for g in groups], glob.data_ptr(), mix.data_ptr(), pm, po2)


run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(a.reps):
    run()
e1.record(stream)
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / a.reps
out = {"circuit": sc.name, "steps": sc.n_steps, "taps": sc.n_taps, "po2": po2, "eval_check_ms": ms,
       "threads": os.environ.get("B2Z_EVALCHECK_THREADS", "128"), "points_per_s": 4 * n / (ms * 1e-3)}
del groups, check
if a.prove:
    zk = 1994
    dcode, ddata = hal.alloc(wc * n), hal.alloc(wd * n)
    dglob, dnoise = hal.alloc(sc.globals_size), hal.alloc(wa * zk)
    hal.witgen(c, po2, dcode, ddata, dglob, dnoise, 0xB2000002, zk)
    g = hal.view(dglob, sc.globals_size)
    hal.prove_segment_device(c, po2, dcode, ddata, g, dnoise, zk)
    e0.record(stream)
    for _ in range(a.reps):
        seal = hal.prove_segment_device(c, po2, dcode, ddata, g, dnoise, zk)
    e1.record(stream)
    torch.cuda.synchronize()
    out["segment_ms"] = e0.elapsed_time(e1) / a.reps
    out["phase_ms"] = hal.last_phase_ms()
    c.set_control_root(po2, hal.compute_control_root(po2, dcode, wc))
    out["verify"] = c.verify(seal)
print(json.dumps(out))
c.close()
hal.close()
