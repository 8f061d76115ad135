#!/usr/bin/env python3
"""bench.py -- RV32IM cycles proved per second on B200 (BASELINE.json metric), one segment proof per step.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--po2 P] [--impl b2z|reference]

A step = one pass of the hot path over one batch = ONE segment proof (iNTT/zk_shift -> x4 LDE NTT -> Poseidon2 Merkle
commits of code/data/accum -> eval_check -> check commit -> DEEP -> FRI -> 50 queries) of the planning-width synthetic
circuit (16 + 192 + 48 columns; the real rv32im circuit is not available offline).  Default workload is
BASELINE.json configs[1]: one po2 = 22 segment per step on one B200.  With N > 1 every rank proves its own segments
(segments are independent: weak scaling, no data-path collective); NCCL only broadcasts the segment table and gathers
the seals (timed separately).

  value  : cycles/s with the witness already resident in HBM (b2z_prove_segment_device), CUDA-event timed, max over ranks
  e2e    : same metric through the host-buffer C-ABI call (b2z_prove_segment): pinned host witness -> H2D -> prove -> seal D2H
  roofline      : the dominant kernel (Poseidon2 hash_rows over the data-group LDE) timed alone with CUDA events
  cpu_baseline  : the CPU restatement (oracle/, "port") on the host cores over a bounded sample -- NOT risc0's binary
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "rv32im_cycles_proved_per_sec"
UNIT = "cycles/s"
SEED0 = 0xB2000000


def planning_circuit():
    from zeth_b200 import synth_circuit as S
    return S.circuit_planning()


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self) -> dict:
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


def usable_cpus() -> int:
    """CPUs this process may really use: affinity mask capped by the cgroup CPU quota (containers often expose every
    hardware thread of the host while limiting CPU time; spinning OpenMP threads beyond the quota slow everything down)."""
    n = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
    try:
        quota, period = open("/sys/fs/cgroup/cpu.max").read().split()
        if quota != "max":
            n = min(n, max(1, int(int(quota) / int(period))))
    except Exception:
        pass
    return max(1, n)


def cpu_reference_run(steps: int, warmup: int, target_s: float = 12.0):
    """Time the CPU restatement (oracle/) on the host cores over bounded samples of the same workload."""
    # all usable cores (torchrun pins OMP_NUM_THREADS=1, which would cripple the CPU arm); no busy-waiting between regions
    os.environ["OMP_NUM_THREADS"] = str(usable_cpus())
    os.environ.setdefault("OMP_WAIT_POLICY", "PASSIVE")
    os.environ.setdefault("OMP_PROC_BIND", "false")
    from oracle import pyoracle as O
    sc = planning_circuit()
    oc = O.Circuit(sc.blob)
    cores = int(O.lib().or_num_threads())

    def one(po2, seed):
        n = 1 << po2
        zk = min(1994, n // 4)
        code, data, glob, noise = oc.witgen(po2, seed, zk)
        t = time.perf_counter()
        seal = oc.prove(po2, code, data, glob, noise, zk)
        dt = time.perf_counter() - t
        assert seal.size > 0
        return dt

    one(12, 1)  # warm the thread pool / page in
    # samples smaller than po2 = 16 are dominated by fixed per-phase overheads on many-core hosts and under-report the CPU
    probe_po2 = int(os.environ.get("B2Z_REF_MIN_PO2", "16"))  # (tests shrink it)
    est = min(one(probe_po2, 2), one(probe_po2, 3))
    total_runs = max(1, steps + warmup)
    per_run = max(target_s / total_runs, 0.5)
    po2 = probe_po2
    while po2 < 20 and est * 2.1 <= per_run:
        po2 += 1
        est *= 2.1
    for w in range(warmup):
        one(po2, 100 + w)
    ts = [one(po2, 200 + k) for k in range(max(1, steps))]
    cycles = (1 << po2) * len(ts)
    value = cycles / sum(ts)
    sample = f"{len(ts)} x po2={po2} segment(s) of the planning circuit (W=256), {sum(ts):.1f} s on {cores} host threads (OpenMP)"
    return value, cores, sample, 1e3 * sum(ts) / len(ts), po2


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    value, cores, sample, ms, po2 = cpu_reference_run(args.steps, args.warmup, target_s=float(os.environ.get("B2Z_REF_TARGET_S", "45")))
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
        "data": "synthetic",
        "config": {"workload": f"segment proof, planning circuit W=256 (16+192+48), bounded CPU sample po2={po2} of the po2={args.po2} workload",
                   "note": "CPU restatement (oracle/), not risc0's binary: risc0's prover cannot be built offline (no Rust, crates not vendored)"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def run_b2z(args):
    import torch
    import torch.distributed as dist
    from zeth_b200 import hal as H
    from zeth_b200 import shard

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the b2z arm has no CPU fallback (use --impl reference for the CPU baseline)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner on stdout at the first collective; keep stdout for the ONE JSON line
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            os.close(saved_stdout)

    sc = planning_circuit()
    po2 = args.po2
    n = 1 << po2
    zk = min(H.ZK_CYCLES, n // 4)
    wa, wc, wd = sc.group_sizes
    K, W = args.steps, args.warmup

    hal = H.Hal(local_rank)
    stream = torch.cuda.Stream(device=dev)
    hal.set_stream(stream.cuda_stream)
    circuit = H.Circuit(hal, sc.blob)

    # segment table: rank 0 decides (index, seed, po2); NCCL broadcast; round-robin partition
    n_segments = world * (K + W)
    table = np.array([[i, SEED0 + 1000 + i, po2] for i in range(n_segments)], dtype=np.int64) if rank == 0 else None
    table = shard.broadcast_segment_table(table, dev)
    mine = shard.partition(n_segments, world, rank)
    assert len(mine) == K + W

    # device-resident witnesses are regenerated per step by the witgen kernels (untimed), one buffer set reused
    with torch.cuda.stream(stream):
        d_code = torch.empty(wc * n, dtype=torch.int32, device=dev)
        d_data = torch.empty(wd * n, dtype=torch.int32, device=dev)
        d_glob = torch.empty(max(sc.globals_size, 1), dtype=torch.int32, device=dev)
        d_noise = torch.empty(max(wa * zk, 1), dtype=torch.int32, device=dev)

    def gen(seg):
        hal.witgen(circuit, po2, d_code.data_ptr(), d_data.data_ptr(), d_glob.data_ptr(), d_noise.data_ptr(), int(table[seg][1]), zk)
        hal.sync()
        return d_glob.cpu().numpy().view(np.uint32)[: sc.globals_size].copy()

    seals = {}
    # ---- warm-up ----
    for seg in mine[:W]:
        g = gen(seg)
        seals[seg] = hal.prove_segment_device(circuit, po2, d_code.data_ptr(), d_data.data_ptr(), g, d_noise.data_ptr(), zk)
    # ---- timed: K segment proofs, witness resident in HBM (each >> L2: 3.3 GiB witness at po2 22; no L2 flush needed) ----
    launches0 = hal.launch_count()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    step_ms = []
    phase_acc = {}
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    total_ms = 0.0
    witgen_launches = 0
    for seg in mine[W:]:
        l0 = hal.launch_count()
        g = gen(seg)  # untimed: the witness producer (f1 row) is outside the metric's region
        witgen_launches += hal.launch_count() - l0
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        seals[seg] = hal.prove_segment_device(circuit, po2, d_code.data_ptr(), d_data.data_ptr(), g, d_noise.data_ptr(), zk)
        e1.record(stream)
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        step_ms.append(ms)
        total_ms += ms
        for k, v in hal.last_phase_ms().items():
            phase_acc[k] = phase_acc.get(k, 0.0) + v
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop() if rank == 0 else None
    launches = hal.launch_count() - launches0 - witgen_launches
    t = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    max_ms = float(t.item())
    value = world * K * n / (max_ms * 1e-3)

    # ---- seal gather over NCCL (timed separately; not part of the metric) ----
    gather_ms = None
    if world > 1:
        shard.gather_seals(seals, n_segments, circuit.seal_words(po2), dev)  # warm the gather channels
        dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        allseals = shard.gather_seals(seals, n_segments, circuit.seal_words(po2), dev)
        torch.cuda.synchronize()
        gather_ms = (time.perf_counter() - t0) * 1e3
        if rank == 0:
            assert len(allseals) == n_segments

    # ---- e2e: host-buffer C-ABI call, pinned host witness, H2D + prove + seal D2H inside the timed region ----
    with torch.cuda.stream(stream):
        h_code = torch.empty(wc * n, dtype=torch.int32).pin_memory()
        h_data = torch.empty(wd * n, dtype=torch.int32).pin_memory()
        h_noise = torch.empty(max(wa * zk, 1), dtype=torch.int32).pin_memory()
    g = gen(mine[-1])
    h_code.copy_(d_code)
    h_data.copy_(d_data)
    h_noise.copy_(d_noise)
    torch.cuda.synchronize()
    seal_buf = np.empty(circuit.seal_words(po2), dtype=np.uint32)
    e2e_steps = max(1, min(K, 3))
    for _ in range(1):
        hal.prove_segment(circuit, po2, h_code.data_ptr(), h_data.data_ptr(), g, h_noise.data_ptr(), zk, seal_buf)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(e2e_steps):
        s = hal.prove_segment(circuit, po2, h_code.data_ptr(), h_data.data_ptr(), g, h_noise.data_ptr(), zk, seal_buf)
    e1.record(stream)
    torch.cuda.synchronize()
    e2e_ms = max(e0.elapsed_time(e1), (time.perf_counter() - t0) * 1e3)
    assert np.array_equal(s, seals[mine[-1]]), "host-buffer path and device-resident path disagree"
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = world * e2e_steps * n / (float(t.item()) * 1e-3)
    h2d = 4 * (wc * n + wd * n + wa * zk) + 4 * sc.globals_size
    d2h = 4 * circuit.seal_words(po2)

    # ---- roofline of the dominant kernel, timed alone with CUDA events on the launching stream ----
    roofline = None
    kernels = []
    cpu_baseline = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        rows = 4 * n

        def time_kernel(fn, reps=5):
            fn()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream)
            for _ in range(reps):
                fn()
            b.record(stream)
            torch.cuda.synchronize()
            return a.elapsed_time(b) / reps

        with torch.cuda.stream(stream):
            lde = torch.empty(wd * rows, dtype=torch.int32, device=dev)
            dig = torch.empty(8 * rows, dtype=torch.int32, device=dev)
        hal.batch_expand_into_evaluate_ntt(lde.data_ptr(), d_data.data_ptr(), wd, po2, 2)
        ms_hr = time_kernel(lambda: hal.hash_rows(dig.data_ptr(), lde.data_ptr(), rows, wd))
        bytes_hr = 4 * wd * rows + 32 * rows
        perms = rows * ((wd + 15) // 16)
        # dram bytes of this kernel from the committed `ncu --set full` capture of the same shape (profiles/), else null
        traffic = None
        try:
            tr = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))["hash_rows_kernel"]
            for ent in tr:
                if ent["rows"] == rows and ent["cols"] == wd:
                    traffic = ent["dram_bytes"]
        except Exception:
            pass
        roofline = {"kernel": "hash_rows_mixed_kernel (Poseidon2 leaves of the data-group LDE; INT + exact-FP64 instruction streams)", "bound": "hbm", "achieved": bytes_hr / ms_hr / 1e6,
                    "peak": hbm, "unit": "GB/s", "frac": bytes_hr / ms_hr / 1e6 / hbm, "traffic": traffic, "peak_source": peak_src,
                    "ms_per_launch": ms_hr, "algorithmic_bytes": bytes_hr,
                    "alu_note": "arithmetic-pipe bound, not HBM bound: 1356 modular products per 64 B absorbed; measured INT32 product "
                                "peak 3.63e12/s (tools/microbench/intpipe.cu) => 2.7e9 permutations/s for the integer stream alone; "
                                "half of the CTAs run the permutation as exact FP64 arithmetic on the FP64 pipe (DESIGN.md 4.0)",
                    "alu_frac_of_product_peak": 1356 * perms / (ms_hr * 1e-3) / 3.63e12,
                    "poseidon2_perms_per_s": perms / (ms_hr * 1e-3), "modmul_per_s": 1356 * perms / (ms_hr * 1e-3)}
        ms_ex = time_kernel(lambda: hal.batch_expand_into_evaluate_ntt(lde.data_ptr(), d_data.data_ptr(), wd, po2, 2))
        bytes_ex = 20 * wd * n
        kernels.append({"kernel": "batch_expand_into_evaluate_ntt (2 passes)", "ms": ms_ex, "algorithmic_bytes": bytes_ex,
                        "achieved_gbs": bytes_ex / ms_ex / 1e6, "frac_hbm": bytes_ex / ms_ex / 1e6 / hbm})
        with torch.cuda.stream(stream):
            tmp = d_data.clone()
        ms_in = time_kernel(lambda: hal.batch_interpolate_ntt(tmp.data_ptr(), wd, po2))
        bytes_in = 8 * wd * n
        kernels.append({"kernel": "batch_interpolate_ntt (2 passes)", "ms": ms_in, "algorithmic_bytes": bytes_in,
                        "achieved_gbs": bytes_in / ms_in / 1e6, "frac_hbm": bytes_in / ms_in / 1e6 / hbm})
        del lde, dig, tmp
        if not args.no_cpu_baseline and world == 1:  # reported at N = 1 only
            v, cores, sample, _, _ = cpu_reference_run(1, 0, target_s=12.0)
            cpu_baseline = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": max_ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u32",
            "data": "synthetic",
            "config": {"workload": f"BASELINE configs[1]: one po2={po2} segment proof per step per GPU, synthetic planning circuit W=256 "
                                   f"(code 16 + data 192 + accum 48), zk rows {zk}", "po2": po2, "cycles_per_step_per_gpu": n,
                       "segments_per_step": world, "parallelism": f"segments round-robin over {world} GPU(s), no data-path collective",
                       "l2": "inputs >> L2 (witness 4*(16+192)*2^po2 B per step), no flush needed",
                       "timing": "CUDA events on the launching stream per step, summed, max over ranks"},
            "clocks": clocks,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "steps": e2e_steps,
                    "api": "b2z_prove_segment (host pinned buffers in, seal out)"},
            "gpu_launches": int(launches),
            "roofline": roofline,
            "kernels": kernels,
            "cpu_baseline": cpu_baseline,
            "phase_ms_per_step": {k: v / K for k, v in phase_acc.items()},
            "seal_gather_ms": gather_ms,
        }
        print(json.dumps(line))
    circuit.close()
    hal.close()
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--po2", type=int, default=22)
    ap.add_argument("--impl", default="b2z", choices=["b2z", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    return run_b2z(args)


if __name__ == "__main__":
    sys.exit(main())
