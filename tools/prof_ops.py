#!/usr/bin/env python3
"""Run individual Hal ops at bench shapes (for targeted ncu captures and CUDA-event timings).
usage: python tools/prof_ops.py [--po2 20] [--cols 192] [--ops expand,interp,hash_rows,hash_fold,bitrev,eval_check] [--reps 3]"""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zeth_b200 import hal as H, synth_circuit as S  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--po2", type=int, default=20)
ap.add_argument("--cols", type=int, default=192)
ap.add_argument("--ops", default="expand,interp,hash_rows,hash_fold,bitrev")
ap.add_argument("--reps", type=int, default=3)
a = ap.parse_args()
dev = torch.device("cuda", 0)
hal = H.Hal(0)
stream = torch.cuda.Stream(device=dev)
hal.set_stream(stream.cuda_stream)
n, cols, po2 = 1 << a.po2, a.cols, a.po2
P = 15 * (1 << 27) + 1
with torch.cuda.stream(stream):
    x = torch.randint(0, P, (cols * n,), dtype=torch.int32, device=dev)
    lde = torch.empty(cols * 4 * n, dtype=torch.int32, device=dev)
    dig = torch.empty(2 * 8 * 4 * n, dtype=torch.int32, device=dev)


def timeit(name, fn, bytes_):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(a.reps):
        fn()
    e1.record(stream)
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.reps
    print(f"{name:28s} {ms:9.3f} ms  {bytes_ / ms / 1e6:9.1f} GB/s (algorithmic)")


ops = a.ops.split(",")
if "expand" in ops:
    timeit("expand_ntt x4", lambda: hal.batch_expand_into_evaluate_ntt(lde.data_ptr(), x.data_ptr(), cols, po2, 2), 20 * cols * n)
if "interp" in ops:
    timeit("interpolate_ntt", lambda: hal.batch_interpolate_ntt(x.data_ptr(), cols, po2), 8 * cols * n)
if "bitrev" in ops:
    timeit("bit_reverse", lambda: hal.batch_bit_reverse(x.data_ptr(), cols, po2), 8 * cols * n)
if "hash_rows" in ops:
    timeit("hash_rows (4n rows)", lambda: hal.hash_rows(dig.data_ptr() + 32 * 4 * n, lde.data_ptr(), 4 * n, cols), 4 * cols * 4 * n + 32 * 4 * n)
if "hash_fold" in ops:
    timeit("hash_fold (4n -> 2n)", lambda: hal.hash_fold(dig.data_ptr(), 4 * n, 2 * n), 96 * 2 * n)
if "merkle" in ops:
    timeit("merkle_build", lambda: hal.merkle_build(dig.data_ptr(), lde.data_ptr(), 4 * n, cols), 4 * cols * 4 * n + 32 * 4 * n)
hal.close()
