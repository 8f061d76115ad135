"""ctypes binding of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module;
the product package `zeth_b200` never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
P = 15 * (1 << 27) + 1


def build(force: bool = False) -> str:
    """liboracle.so (x86-64-v3, AVX2) and liboracle_v4.so (x86-64-v4, AVX-512) from the same sources"""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("bb.c", "poseidon2.c", "hal.c", "circuit.c", "zkp.c", "bb.h", "oracle.h", "poseidon2_consts.h",
                                              "../include/protocol_params.h")]
    for name in ("liboracle.so", "liboracle_v4.so"):
        out = os.path.join(_HERE, name)
        if force or not os.path.exists(out) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in srcs):
            subprocess.check_call(["make", "-C", _HERE, "-s", name] + (["-B"] if force else []))
    return so


def isa() -> str:
    """Which build lib() loads on this host: 'x86-64-v4' when the CPU has the AVX-512 subsets of that level, else 'x86-64-v3'."""
    if os.environ.get("B2Z_ORACLE_ISA", "") == "v3":
        return "x86-64-v3"
    try:
        flags = set()
        for line in open("/proc/cpuinfo"):
            if line.startswith("flags"):
                flags = set(line.split(":", 1)[1].split())
                break
        if {"avx512f", "avx512dq", "avx512bw", "avx512vl", "avx512cd"} <= flags:
            return "x86-64-v4"
    except OSError:
        pass
    return "x86-64-v3"


def lib():
    global _LIB
    if _LIB is None:
        name = "liboracle_v4.so" if isa() == "x86-64-v4" else "liboracle.so"
        so = os.path.join(_HERE, name)
        if not os.path.exists(so):
            build()
        _LIB = C.CDLL(so)
        L = _LIB
        L.or_circuit_load.restype = C.c_void_p
        L.or_circuit_load.argtypes = [C.c_void_p, C.c_size_t]
        L.or_circuit_free.argtypes = [C.c_void_p]
        L.or_prove_segment.restype = C.c_size_t
        L.or_prove_segment.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_size_t,
                                       C.POINTER(C.POINTER(C.c_uint32)), C.c_void_p]
        L.or_verify_segment.restype = C.c_int
        L.or_verify_segment.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        L.or_free.argtypes = [C.c_void_p]
        L.or_num_threads.restype = C.c_int
        L.or_synth_rand.restype = C.c_uint32
        L.or_synth_rand.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint64]
        L.or_rng_elem.restype = C.c_uint32
        L.or_rng_bits.restype = C.c_uint32
    return _LIB


def _p(a: np.ndarray):
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def u32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint32)


def to_mont(a) -> np.ndarray:
    a = u32(a).copy()
    lib().or_to_montgomery(_p(a), C.c_size_t(a.size))
    return a


def from_mont(a) -> np.ndarray:
    a = u32(a).copy()
    lib().or_from_montgomery(_p(a), C.c_size_t(a.size))
    return a


def fp_mul(a, b):
    a, b = u32(a), u32(b)
    out = np.empty_like(a)
    lib().or_fp_mul_vec(_p(out), _p(a), _p(b), C.c_size_t(a.size))
    return out


def fp4_mul(a, b):
    a, b = u32(a), u32(b)
    out = np.empty_like(a)
    lib().or_fp4_mul_vec(_p(out), _p(a), _p(b), C.c_size_t(a.size // 4))
    return out


def fp4_inv(a):
    a = u32(a)
    out = np.empty_like(a)
    lib().or_fp4_inv_vec(_p(out), _p(a), C.c_size_t(a.size // 4))
    return out


def poseidon2_mix(state) -> np.ndarray:
    s = u32(state).copy()
    assert s.size == 24
    lib().or_poseidon2_mix(_p(s))
    return s


def hash_elem_slice(elems) -> np.ndarray:
    e = u32(elems)
    out = np.empty(8, dtype=np.uint32)
    lib().or_hash_elem_slice(_p(e), C.c_size_t(e.size), _p(out))
    return out


def hash_pair(a, b) -> np.ndarray:
    a, b = u32(a), u32(b)
    out = np.empty(8, dtype=np.uint32)
    lib().or_hash_pair(_p(a), _p(b), _p(out))
    return out


class Rng:
    """Poseidon2Rng (risc0-zkp core/hash/poseidon2/rng.rs)."""

    def __init__(self):
        self.buf = np.zeros(25, dtype=np.uint32)
        lib().or_rng_init(_p(self.buf))

    def mix(self, digest):
        d = u32(digest)
        lib().or_rng_mix(_p(self.buf), _p(d))

    def random_elem(self) -> int:
        return int(lib().or_rng_elem(_p(self.buf)))

    def random_bits(self, bits: int) -> int:
        return int(lib().or_rng_bits(_p(self.buf), C.c_uint(bits)))


# ---- Hal ops (host memory, column-major [count][n]) ----
def batch_interpolate_ntt(io, count):
    io = u32(io).copy()
    lib().or_batch_interpolate_ntt(_p(io), C.c_size_t(count), C.c_size_t(io.size // count))
    return io


def batch_evaluate_ntt(io, count, expand_bits=0):
    io = u32(io).copy()
    lib().or_batch_evaluate_ntt(_p(io), C.c_size_t(count), C.c_size_t(io.size // count), C.c_uint(expand_bits))
    return io


def batch_expand_into_evaluate_ntt(inp, count, expand_bits=2):
    inp = u32(inp)
    n = inp.size // count
    out = np.empty(count * (n << expand_bits), dtype=np.uint32)
    lib().or_batch_expand_into_evaluate_ntt(_p(out), _p(inp), C.c_size_t(count), C.c_size_t(n), C.c_uint(expand_bits))
    return out


def batch_bit_reverse(io, count):
    io = u32(io).copy()
    lib().or_batch_bit_reverse(_p(io), C.c_size_t(count), C.c_size_t(io.size // count))
    return io


def zk_shift(io, count):
    io = u32(io).copy()
    lib().or_zk_shift(_p(io), C.c_size_t(count), C.c_size_t(io.size // count))
    return io


def hash_rows(matrix, rows, cols):
    m = u32(matrix)
    out = np.empty(rows * 8, dtype=np.uint32)
    lib().or_hash_rows(_p(out), _p(m), C.c_size_t(rows), C.c_size_t(cols))
    return out


def hash_fold(nodes, in_size, out_size):
    nodes = u32(nodes).copy()
    lib().or_hash_fold(_p(nodes), C.c_size_t(in_size), C.c_size_t(out_size))
    return nodes


def merkle_nodes(matrix, rows, cols):
    """Full node array (2*rows digests) as MerkleTreeProver::new builds it."""
    nodes = np.zeros(2 * rows * 8, dtype=np.uint32)
    nodes[8 * rows:] = hash_rows(matrix, rows, cols)
    size = rows
    while size > 1:
        size //= 2
        nodes = hash_fold(nodes, size * 2, size)
    return nodes


def fri_fold(inp, mix):
    inp, mix = u32(inp), u32(mix)
    count = inp.size // 4 // 16
    out = np.empty(4 * count, dtype=np.uint32)
    lib().or_fri_fold(_p(out), _p(inp), _p(mix), C.c_size_t(count))
    return out


def batch_evaluate_any(coeffs, poly_count, which, xs):
    coeffs, which, xs = u32(coeffs), u32(which), u32(xs)
    n = coeffs.size // poly_count
    out = np.empty(4 * which.size, dtype=np.uint32)
    lib().or_batch_evaluate_any(_p(coeffs), C.c_size_t(poly_count), C.c_size_t(n), _p(which), _p(xs), C.c_size_t(which.size), _p(out))
    return out


def mix_poly_coeffs(out, mix_start, mix, inp, combos, count):
    out = u32(out).copy()
    mix_start, mix, inp, combos = u32(mix_start), u32(mix), u32(inp), u32(combos)
    lib().or_mix_poly_coeffs(_p(out), _p(mix_start), _p(mix), _p(inp), _p(combos), C.c_size_t(combos.size), C.c_size_t(count))
    return out


def eltwise_sum_extelem(inp, count):
    inp = u32(inp)
    to_add = inp.size // 4 // count
    out = np.empty(4 * count, dtype=np.uint32)
    lib().or_eltwise_sum_extelem(_p(out), _p(inp), C.c_size_t(count), C.c_size_t(to_add))
    return out


def scatter(into, index, offsets, values):
    into = u32(into).copy()
    index, offsets, values = u32(index), u32(offsets), u32(values)
    lib().or_scatter(_p(into), _p(index), _p(offsets), _p(values), C.c_size_t(index.size - 1))
    return into


def prefix_products(io):
    io = u32(io).copy()
    lib().or_prefix_products(_p(io), C.c_size_t(io.size // 4))
    return io


def control_root(po2, code, cols) -> np.ndarray:
    """Merkle root of the committed code group (risc0: the control ID of this po2)."""
    code = u32(code)
    out = np.empty(8, dtype=np.uint32)
    L = lib()
    L.or_control_root.argtypes = [C.c_uint, C.c_void_p, C.c_size_t, C.c_void_p]
    L.or_control_root(po2, _p(code), cols, _p(out))
    return out


def eltwise_zeroize_elem(io):
    io = u32(io).copy()
    lib().or_eltwise_zeroize_elem(_p(io), C.c_size_t(io.size))
    return io


class Circuit:
    def __init__(self, blob: np.ndarray):
        self.blob = u32(blob)
        self.h = lib().or_circuit_load(_p(self.blob), C.c_size_t(self.blob.size))
        if not self.h:
            raise ValueError("oracle: bad circuit blob")
        L = lib()
        L.or_circuit_group_size.restype = C.c_size_t
        L.or_circuit_group_size.argtypes = [C.c_void_p, C.c_int]
        L.or_circuit_globals_size.restype = C.c_size_t
        L.or_circuit_globals_size.argtypes = [C.c_void_p]
        L.or_circuit_mix_size.restype = C.c_size_t
        L.or_circuit_mix_size.argtypes = [C.c_void_p]
        self.group_sizes = tuple(int(L.or_circuit_group_size(self.h, g)) for g in range(3))
        self.globals_size = int(L.or_circuit_globals_size(self.h))
        self.mix_size = int(L.or_circuit_mix_size(self.h))

    def __del__(self):
        if getattr(self, "h", None):
            lib().or_circuit_free(self.h)
            self.h = None

    def witgen(self, po2: int, seed: int, zk_rows: int):
        n = 1 << po2
        wa, wc, wd = self.group_sizes
        code = np.empty(wc * n, dtype=np.uint32)
        data = np.empty(wd * n, dtype=np.uint32)
        glob = np.empty(max(self.globals_size, 1), dtype=np.uint32)
        L = lib()
        L.or_witgen.argtypes = [C.c_void_p, C.c_uint, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_size_t]
        L.or_witgen(self.h, po2, _p(code), _p(data), _p(glob), seed, zk_rows)
        # witgen emits the CANONICAL code columns of (po2, zk_rows): register their Merkle root as this po2's control root
        # (risc0 ships that table with the circuit; the verifier rejects seals committing to any other code columns)
        self.set_control_root(po2, control_root(po2, code, wc))
        noise = np.array([L.or_synth_rand(seed, 4, c, r) for c in range(wa) for r in range(zk_rows)], dtype=np.uint32) \
            if wa * zk_rows <= 4096 else synth_rand_array(seed, 4, wa, zk_rows)
        return code, data, glob[: self.globals_size], noise

    def set_control_root(self, po2, root):
        L = lib()
        L.or_circuit_set_control_root.argtypes = [C.c_void_p, C.c_uint, C.c_void_p]
        r = u32(root)
        L.or_circuit_set_control_root(self.h, po2, _p(r))
        self.control_roots = getattr(self, "control_roots", {})
        self.control_roots[po2] = r.copy()

    def accumulate(self, po2, code, data, mix, noise, zk_rows):
        n = 1 << po2
        wa = self.group_sizes[0]
        accum = np.empty(wa * n, dtype=np.uint32)
        L = lib()
        L.or_accumulate.argtypes = [C.c_void_p, C.c_uint] + [C.c_void_p] * 5 + [C.c_size_t]
        L.or_accumulate(self.h, po2, _p(accum), _p(u32(code)), _p(u32(data)), _p(u32(mix)), _p(u32(noise)), zk_rows)
        return accum

    def eval_check(self, po2, groups, globals_, mix, poly_mix, scalar=False):
        """scalar=True: the one-point-at-a-time definition (or_eval_check_scalar); default: the lane-blocked typed form"""
        n = 1 << po2
        check = np.empty(16 * n, dtype=np.uint32)
        gs = [u32(g) for g in groups]
        arr = (C.c_void_p * 3)(*[g.ctypes.data for g in gs])
        fn = lib().or_eval_check_scalar if scalar else lib().or_eval_check
        fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint]
        fn.restype = None
        fn(self.h, _p(check), arr, _p(u32(globals_)), _p(u32(mix)), _p(u32(poly_mix)), po2)
        return check

    def prove(self, po2, code, data, globals_, accum_noise, zk_rows, timings=None):
        seal = C.POINTER(C.c_uint32)()
        tm = (C.c_double * 16)()
        n = lib().or_prove_segment(self.h, po2, _p(u32(code)), _p(u32(data)), _p(u32(globals_)), _p(u32(accum_noise)), zk_rows,
                                   C.byref(seal), C.cast(tm, C.c_void_p))
        out = np.ctypeslib.as_array(seal, shape=(n,)).copy()
        lib().or_free(seal)
        if timings is not None:
            timings.extend(list(tm))
        return out

    def verify(self, seal) -> int:
        s = u32(seal)
        return int(lib().or_verify_segment(self.h, _p(s), C.c_size_t(s.size)))


def synth_rand_array(seed, stream, cols, rows):
    """Vectorised numpy twin of or_synth_rand (splitmix64 counter stream), shape [cols*rows] column-major."""
    def sm(x):
        x = (x + np.uint64(0x9E3779B97F4A7C15))
        x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return x ^ (x >> np.uint64(31))
    with np.errstate(over="ignore"):
        c = np.arange(cols, dtype=np.uint64)[:, None]
        r = np.arange(rows, dtype=np.uint64)[None, :]
        inner = sm((np.uint64(stream) << np.uint64(56)) ^ (c << np.uint64(32)) ^ r)
        z = sm(np.uint64(seed) ^ inner)
    return (z % np.uint64(P)).astype(np.uint32).reshape(-1)
