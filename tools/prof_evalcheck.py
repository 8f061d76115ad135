#!/usr/bin/env python3
"""Time the generated eval_check kernel alone (planning circuit) on random LDE inputs."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from zeth_b200 import hal as H, synth_circuit as S
po2 = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda", 0)
hal = H.Hal(0)
stream = torch.cuda.Stream(device=dev); hal.set_stream(stream.cuda_stream)
sc = S.circuit_planning(); c = H.Circuit(hal, sc.blob)
n = 1 << po2; P = 15 * (1 << 27) + 1
with torch.cuda.stream(stream):
    gs = [torch.randint(0, P, (w * 4 * n,), dtype=torch.int32, device=dev) for w in sc.group_sizes]
    glob = torch.randint(0, P, (sc.globals_size,), dtype=torch.int32, device=dev)
    mix = torch.randint(0, P, (sc.mix_size,), dtype=torch.int32, device=dev)
    check = torch.empty(16 * n, dtype=torch.int32, device=dev)
pm = np.array([5, 6, 7, 8], dtype=np.uint32)
f = lambda: hal.eval_check(c, check.data_ptr(), [g.data_ptr() for g in gs], glob.data_ptr(), mix.data_ptr(), pm, po2)
f(); torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(stream)
for _ in range(3): f()
e1.record(stream); torch.cuda.synchronize()
print(f"eval_check po2={po2} points_per_thread={os.environ.get('B2Z_EVALCHECK_POINTS','1')} "
      f"minblocks={os.environ.get('B2Z_EVALCHECK_MINBLOCKS','default')}: {e0.elapsed_time(e1)/3:.3f} ms")
