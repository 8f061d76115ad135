"""zeth_b200 -- B200-native (sm_100a) STARK segment prover for the RISC Zero zkVM path that risc0/zeth drives through
`default_prover().prove(env, elf)` (/root/reference crates/host/src/lib.rs:137).

The product is `libb2z.so` (hand-written CUDA kernels + a C++ prover driver behind the C ABI of include/b2z.h).
This Python package is only the host-side mirror used by tests, bench.py and multi-GPU plumbing:
  zeth_b200.hal       ctypes mirror of the upstream `Hal` / `CircuitHal` / segment-prover surface
  zeth_b200.synth_circuit   synthetic circuits (the real rv32im circuit is not available offline)
  zeth_b200.shard     one-process-per-GPU segment sharding over torch.distributed
There is no CPU fallback: every compute call fails loudly if the CUDA extension or a B200 is missing.
"""
from .hal import B2ZError, Circuit, Hal, lib_path, load_library  # noqa: F401
