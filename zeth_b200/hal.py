"""ctypes mirror of the C ABI in include/b2z.h: the upstream `Hal` trait (risc0-zkp 3.0.2 hal/mod.rs), `CircuitHal`
(risc0-circuit-rv32im 4.0.2) and `SegmentProver::prove` surface, same method names and argument meaning.

Buffers are raw device pointers (ints) so that torch tensors (`t.data_ptr()`) or library allocations both work.
Errors: any non-NULL return of the C ABI raises B2ZError (upstream: anyhow::Error from the FFI string).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

INV_RATE, QUERIES, FRI_FOLD, FRI_MIN_DEGREE, EXT_SIZE, CHECK_SIZE, DIGEST_WORDS, ZK_CYCLES = 4, 50, 16, 256, 4, 16, 8, 1994
GROUP_ACCUM, GROUP_CODE, GROUP_DATA = 0, 1, 2


class B2ZError(RuntimeError):
    pass


def lib_path() -> str:
    return os.path.join(_HERE, "libb2z.so")


# every symbol include/b2z.h declares (checked by tests/test_abi.py against the header text)
_ERR_FUNCS = [
    "b2z_init", "b2z_destroy", "b2z_set_stream", "b2z_sync", "b2z_alloc", "b2z_free", "b2z_h2d", "b2z_d2h",
    "b2z_batch_interpolate_ntt", "b2z_batch_evaluate_ntt", "b2z_batch_expand_into_evaluate_ntt", "b2z_batch_bit_reverse",
    "b2z_zk_shift", "b2z_hash_rows", "b2z_hash_rows_chunk", "b2z_hash_fold", "b2z_merkle_build", "b2z_fri_fold", "b2z_batch_evaluate_any",
    "b2z_mix_poly_coeffs", "b2z_eltwise_add_elem", "b2z_eltwise_sum_extelem", "b2z_eltwise_copy_elem",
    "b2z_eltwise_zeroize_elem", "b2z_gather_sample", "b2z_poly_divide", "b2z_circuit_load", "b2z_circuit_precompile",
    "b2z_circuit_free", "b2z_eval_check", "b2z_accumulate", "b2z_witgen", "b2z_prove_segment_device", "b2z_prove_segment",
    "b2z_scatter", "b2z_prefix_products", "b2z_compute_control_root", "b2z_circuit_set_control_root",
    "b2z_session_create", "b2z_session_destroy", "b2z_session_set_control_root", "b2z_session_set_seal_cache", "b2z_prove_session",
    "b2z_receipt_assemble", "b2z_receipt_free", "b2z_receipt_serialize", "b2z_receipt_deserialize",
]


def load_library():
    """Load libb2z.so; raises if it has not been built (python -m zeth_b200.build / __graft_entry__.build())."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise B2ZError(f"{path} is missing: build the CUDA extension first (__graft_entry__.build()); there is no CPU fallback")
    L = C.CDLL(path)
    for name in _ERR_FUNCS:
        getattr(L, name).restype = C.c_char_p
    L.b2z_version.restype = C.c_char_p
    L.b2z_launch_count.restype = C.c_uint64
    L.b2z_launch_count.argtypes = [C.c_void_p]
    for name in ("b2z_circuit_group_size", "b2z_circuit_globals_size", "b2z_circuit_mix_size", "b2z_circuit_seal_words", "b2z_last_phase_ms"):
        getattr(L, name).restype = C.c_size_t
    L.b2z_circuit_source.restype = C.c_char_p
    L.b2z_phase_name.restype = C.c_char_p
    L.b2z_phase_name.argtypes = [C.c_size_t]
    L.b2z_verify_segment.restype = C.c_int
    L.b2z_verify_receipt.restype = C.c_int
    L.b2z_circuit_get_control_root.restype = C.c_int
    L.b2z_session_circuit.restype = C.c_void_p
    L.b2z_session_circuit.argtypes = [C.c_void_p]
    L.b2z_session_context.restype = C.c_void_p
    L.b2z_session_context.argtypes = [C.c_void_p, C.c_uint]
    L.b2z_session_in_flight.restype = C.c_uint
    L.b2z_session_in_flight.argtypes = [C.c_void_p]
    L.b2z_session_launch_count.restype = C.c_uint64
    L.b2z_session_launch_count.argtypes = [C.c_void_p]
    L.b2z_session_last_wall_ms.restype = C.c_double
    L.b2z_session_last_wall_ms.argtypes = [C.c_void_p]
    L.b2z_session_last_retries.restype = C.c_size_t
    L.b2z_session_last_retries.argtypes = [C.c_void_p]
    L.b2z_session_last_segment_ms.restype = C.c_size_t
    L.b2z_session_last_cache_hits.restype = C.c_size_t
    L.b2z_session_last_cache_hits.argtypes = [C.c_void_p]
    L.b2z_bytes_free.restype = None
    L.b2z_bytes_free.argtypes = [C.c_void_p]
    _LIB = L
    return L


def _chk(err):
    if err:
        raise B2ZError(err.decode())


def _vp(x):
    if isinstance(x, np.ndarray):
        assert x.flags["C_CONTIGUOUS"]
        return C.c_void_p(x.ctypes.data)
    return C.c_void_p(int(x))


def _u32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.uint32)


class Circuit:
    """A loaded circuit (taps + constraint program + generated kernels)."""

    def __init__(self, hal: "Hal | None", blob):
        L = load_library()
        self.blob = _u32(blob)
        self.h = C.c_void_p()
        _chk(L.b2z_circuit_load(hal.ctx if hal else None, _vp(self.blob), C.c_size_t(self.blob.size), C.byref(self.h)))
        self.group_sizes = tuple(int(L.b2z_circuit_group_size(self.h, C.c_int(g))) for g in range(3))
        self.globals_size = int(L.b2z_circuit_globals_size(self.h))
        self.mix_size = int(L.b2z_circuit_mix_size(self.h))

    def seal_words(self, po2: int) -> int:
        return int(load_library().b2z_circuit_seal_words(self.h, C.c_uint(po2)))

    def source(self) -> str:
        return load_library().b2z_circuit_source(self.h).decode()

    def close(self):
        if self.h:
            load_library().b2z_circuit_free(self.h)
            self.h = C.c_void_p()

    def verify(self, seal) -> int:
        """Host-side STARK verifier (receipt.verify for one segment): 0 = accepted, else a rejection code
        (14: code root != registered control root, 15: no control root registered for the seal's po2)."""
        s = _u32(seal)
        L = load_library()
        return int(L.b2z_verify_segment(self.h, _vp(s), C.c_size_t(s.size)))

    def set_control_root(self, po2: int, root):
        """Register the expected Merkle root of the code group for `po2` (risc0: the control-ID table entry)."""
        r = _u32(root)
        assert r.size == 8
        _chk(load_library().b2z_circuit_set_control_root(self.h, C.c_uint(po2), _vp(r)))

    def control_root(self, po2: int):
        r = np.empty(8, dtype=np.uint32)
        return r if load_library().b2z_circuit_get_control_root(self.h, C.c_uint(po2), _vp(r)) else None

    def verify_receipt(self, receipt: "Receipt", image_id=None):
        """Receipt::verify(image_id) (/root/reference crates/host/src/bin/cli.rs:103): returns (code, bad_segment)."""
        bad = C.c_size_t()
        iid = _vp(_u32(image_id)) if image_id is not None else None
        rc = int(load_library().b2z_verify_receipt(self.h, receipt.ptr, iid, C.byref(bad)))
        return rc, int(bad.value)

    @staticmethod
    def precompile(blob):
        blob = _u32(blob)
        _chk(load_library().b2z_circuit_precompile(_vp(blob), C.c_size_t(blob.size)))


class Hal:
    """One context on one GPU (upstream: CudaHal + circuit hal).  Method names follow `trait Hal`."""

    def __init__(self, device: int = 0):
        L = load_library()
        self.L = L
        self.ctx = C.c_void_p()
        _chk(L.b2z_init(C.c_int(device), C.byref(self.ctx)))
        self.device = device

    def close(self):
        if self.ctx:
            self.L.b2z_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- memory ----
    def set_stream(self, cuda_stream: int):
        _chk(self.L.b2z_set_stream(self.ctx, C.c_void_p(cuda_stream)))

    def sync(self):
        _chk(self.L.b2z_sync(self.ctx))

    def alloc(self, words: int) -> int:
        p = C.c_void_p()
        _chk(self.L.b2z_alloc(self.ctx, C.c_size_t(4 * words), C.byref(p)))
        return p.value

    def free(self, dev: int):
        _chk(self.L.b2z_free(self.ctx, C.c_void_p(dev)))

    def copy_from(self, host) -> int:
        host = _u32(host)
        dev = self.alloc(max(host.size, 1))
        _chk(self.L.b2z_h2d(self.ctx, C.c_void_p(dev), _vp(host), C.c_size_t(4 * host.size)))
        self.sync()
        return dev

    def view(self, dev: int, words: int) -> np.ndarray:
        out = np.empty(words, dtype=np.uint32)
        _chk(self.L.b2z_d2h(self.ctx, _vp(out), C.c_void_p(dev), C.c_size_t(4 * words)))
        return out

    def launch_count(self) -> int:
        return int(self.L.b2z_launch_count(self.ctx))

    # ---- trait Hal ----
    def batch_interpolate_ntt(self, io: int, count: int, po2: int):
        _chk(self.L.b2z_batch_interpolate_ntt(self.ctx, C.c_void_p(io), C.c_size_t(count), C.c_uint(po2)))

    def batch_evaluate_ntt(self, io: int, count: int, po2: int, expand_bits: int = 0):
        _chk(self.L.b2z_batch_evaluate_ntt(self.ctx, C.c_void_p(io), C.c_size_t(count), C.c_uint(po2), C.c_uint(expand_bits)))

    def batch_expand_into_evaluate_ntt(self, out: int, inp: int, count: int, in_po2: int, expand_bits: int = 2):
        _chk(self.L.b2z_batch_expand_into_evaluate_ntt(self.ctx, C.c_void_p(out), C.c_void_p(inp), C.c_size_t(count), C.c_uint(in_po2),
                                                       C.c_uint(expand_bits)))

    def batch_bit_reverse(self, io: int, count: int, po2: int):
        _chk(self.L.b2z_batch_bit_reverse(self.ctx, C.c_void_p(io), C.c_size_t(count), C.c_uint(po2)))

    def zk_shift(self, io: int, count: int, po2: int):
        _chk(self.L.b2z_zk_shift(self.ctx, C.c_void_p(io), C.c_size_t(count), C.c_uint(po2)))

    def hash_rows(self, out: int, matrix: int, rows: int, cols: int):
        _chk(self.L.b2z_hash_rows(self.ctx, C.c_void_p(out), C.c_void_p(matrix), C.c_size_t(rows), C.c_size_t(cols)))

    def hash_rows_chunk(self, out: int, matrix_chunk: int, rows: int, cols: int, carry_in: bool, last: bool):
        _chk(self.L.b2z_hash_rows_chunk(self.ctx, C.c_void_p(out), C.c_void_p(matrix_chunk), C.c_size_t(rows), C.c_size_t(cols),
                                        C.c_int(int(carry_in)), C.c_int(int(last))))

    def hash_fold(self, nodes: int, in_size: int, out_size: int):
        _chk(self.L.b2z_hash_fold(self.ctx, C.c_void_p(nodes), C.c_size_t(in_size), C.c_size_t(out_size)))

    def merkle_build(self, nodes: int, matrix: int, rows: int, cols: int):
        _chk(self.L.b2z_merkle_build(self.ctx, C.c_void_p(nodes), C.c_void_p(matrix), C.c_size_t(rows), C.c_size_t(cols)))

    def fri_fold(self, out: int, inp: int, mix, out_count: int):
        mix = _u32(mix)
        _chk(self.L.b2z_fri_fold(self.ctx, C.c_void_p(out), C.c_void_p(inp), _vp(mix), C.c_size_t(out_count)))

    def batch_evaluate_any(self, coeffs: int, poly_count: int, po2: int, which: int, xs: int, eval_count: int, out: int):
        _chk(self.L.b2z_batch_evaluate_any(self.ctx, C.c_void_p(coeffs), C.c_size_t(poly_count), C.c_uint(po2), C.c_void_p(which),
                                           C.c_void_p(xs), C.c_size_t(eval_count), C.c_void_p(out)))

    def mix_poly_coeffs(self, out: int, mix_start, mix, inp: int, combos: int, input_size: int, count: int):
        ms, m = _u32(mix_start), _u32(mix)
        _chk(self.L.b2z_mix_poly_coeffs(self.ctx, C.c_void_p(out), _vp(ms), _vp(m), C.c_void_p(inp), C.c_void_p(combos),
                                        C.c_size_t(input_size), C.c_size_t(count)))

    def eltwise_add_elem(self, out: int, a: int, b: int, n: int):
        _chk(self.L.b2z_eltwise_add_elem(self.ctx, C.c_void_p(out), C.c_void_p(a), C.c_void_p(b), C.c_size_t(n)))

    def eltwise_sum_extelem(self, out: int, inp: int, count: int, to_add: int):
        _chk(self.L.b2z_eltwise_sum_extelem(self.ctx, C.c_void_p(out), C.c_void_p(inp), C.c_size_t(count), C.c_size_t(to_add)))

    def eltwise_copy_elem(self, out: int, inp: int, n: int):
        _chk(self.L.b2z_eltwise_copy_elem(self.ctx, C.c_void_p(out), C.c_void_p(inp), C.c_size_t(n)))

    def eltwise_zeroize_elem(self, io: int, n: int):
        _chk(self.L.b2z_eltwise_zeroize_elem(self.ctx, C.c_void_p(io), C.c_size_t(n)))

    def gather_sample(self, dst: int, src: int, idx: int, size: int, stride: int):
        _chk(self.L.b2z_gather_sample(self.ctx, C.c_void_p(dst), C.c_void_p(src), C.c_size_t(idx), C.c_size_t(size), C.c_size_t(stride)))

    def scatter(self, into: int, index: int, offsets: int, values: int, count: int):
        _chk(self.L.b2z_scatter(self.ctx, C.c_void_p(into), C.c_void_p(index), C.c_void_p(offsets), C.c_void_p(values), C.c_size_t(count)))

    def prefix_products(self, ext_io: int, n: int):
        _chk(self.L.b2z_prefix_products(self.ctx, C.c_void_p(ext_io), C.c_size_t(n)))

    def compute_control_root(self, po2: int, code_dev: int, code_cols: int) -> np.ndarray:
        root = np.empty(8, dtype=np.uint32)
        _chk(self.L.b2z_compute_control_root(self.ctx, C.c_uint(po2), C.c_void_p(code_dev), C.c_size_t(code_cols), _vp(root)))
        return root

    def poly_divide(self, ext_coeffs: int, count: int, z) -> np.ndarray:
        z = _u32(z)
        rem = np.empty(4, dtype=np.uint32)
        _chk(self.L.b2z_poly_divide(self.ctx, C.c_void_p(ext_coeffs), C.c_size_t(count), _vp(z), _vp(rem)))
        return rem

    # ---- CircuitHal ----
    def eval_check(self, circuit: Circuit, check: int, groups, globals_dev: int, mix_dev: int, poly_mix, po2: int):
        arr = (C.c_void_p * 3)(*[C.c_void_p(int(g)) for g in groups])
        pm = _u32(poly_mix)
        _chk(self.L.b2z_eval_check(self.ctx, circuit.h, C.c_void_p(check), arr, C.c_void_p(globals_dev), C.c_void_p(mix_dev), _vp(pm),
                                   C.c_uint(po2)))

    def accumulate(self, circuit: Circuit, po2: int, accum: int, code: int, data: int, mix_dev: int, noise_dev: int, zk_rows: int):
        _chk(self.L.b2z_accumulate(self.ctx, circuit.h, C.c_uint(po2), C.c_void_p(accum), C.c_void_p(code), C.c_void_p(data),
                                   C.c_void_p(mix_dev), C.c_void_p(noise_dev), C.c_size_t(zk_rows)))

    def witgen(self, circuit: Circuit, po2: int, code: int, data: int, globals_dev: int, noise_dev: int, seed: int, zk_rows: int):
        _chk(self.L.b2z_witgen(self.ctx, circuit.h, C.c_uint(po2), C.c_void_p(code), C.c_void_p(data), C.c_void_p(globals_dev),
                               C.c_void_p(noise_dev), C.c_uint64(seed), C.c_size_t(zk_rows)))

    # ---- SegmentProver::prove ----
    def prove_segment_device(self, circuit: Circuit, po2: int, code_dev: int, data_dev: int, globals_host, noise_dev: int,
                             zk_rows: int) -> np.ndarray:
        g = _u32(globals_host)
        seal = np.empty(circuit.seal_words(po2), dtype=np.uint32)
        n = C.c_size_t()
        _chk(self.L.b2z_prove_segment_device(self.ctx, circuit.h, C.c_uint(po2), C.c_void_p(code_dev), C.c_void_p(data_dev), _vp(g),
                                             C.c_void_p(noise_dev), C.c_size_t(zk_rows), _vp(seal), C.byref(n)))
        return seal[: n.value]

    def prove_segment(self, circuit: Circuit, po2: int, code_host, data_host, globals_host, noise_host, zk_rows: int,
                      seal_out: np.ndarray | None = None) -> np.ndarray:
        """Host buffers in (numpy or pinned torch tensors exposing .ctypes / data_ptr), seal out."""
        g = _u32(globals_host)
        seal = seal_out if seal_out is not None else np.empty(circuit.seal_words(po2), dtype=np.uint32)
        n = C.c_size_t()

        def hp(x):
            return _vp(x) if isinstance(x, np.ndarray) else C.c_void_p(int(x))
        _chk(self.L.b2z_prove_segment(self.ctx, circuit.h, C.c_uint(po2), hp(code_host), hp(data_host), _vp(g), hp(noise_host),
                                      C.c_size_t(zk_rows), _vp(seal), C.byref(n)))
        return seal[: n.value]

    def last_phase_ms(self) -> dict:
        buf = (C.c_double * 32)()
        n = int(self.L.b2z_last_phase_ms(self.ctx, buf, C.c_size_t(32)))
        return {self.L.b2z_phase_name(i).decode(): buf[i] for i in range(min(n, 32))}


# ---- session level: Prover::prove -> Receipt (risc0-zkvm; /root/reference crates/host/src/lib.rs:131-142) ----
class SegmentInput(C.Structure):
    _fields_ = [("index", C.c_uint32), ("po2", C.c_uint32), ("code", C.c_void_p), ("data", C.c_void_p), ("globals", C.c_void_p),
                ("accum_noise", C.c_void_p), ("zk_rows", C.c_size_t), ("on_device", C.c_int)]


class SegmentReceiptStruct(C.Structure):
    _fields_ = [("index", C.c_uint32), ("po2", C.c_uint32), ("hashfn", C.c_char * 16), ("verifier_parameters", C.c_uint32 * 8),
                ("claim_pre", C.c_uint32 * 8), ("claim_post", C.c_uint32 * 8), ("seal", C.POINTER(C.c_uint32)), ("seal_words", C.c_size_t)]


class ReceiptStruct(C.Structure):
    _fields_ = [("segments", C.POINTER(SegmentReceiptStruct)), ("n_segments", C.c_size_t), ("verifier_parameters", C.c_uint32 * 8),
                ("journal", C.POINTER(C.c_uint8)), ("journal_len", C.c_size_t)]


class Receipt:
    """Owner of a b2z_receipt (Receipt{inner: Composite{segments}, journal}); `segments[i]` exposes the C fields."""

    def __init__(self, ptr):
        self.ptr = ptr  # POINTER(ReceiptStruct)

    @property
    def n_segments(self) -> int:
        return int(self.ptr.contents.n_segments)

    def segment(self, i: int) -> SegmentReceiptStruct:
        assert 0 <= i < self.n_segments
        return self.ptr.contents.segments[i]

    def seal(self, i: int) -> np.ndarray:
        s = self.segment(i)
        return np.ctypeslib.as_array(s.seal, shape=(int(s.seal_words),)).copy()

    def claim(self, i: int):
        s = self.segment(i)
        return np.array(list(s.claim_pre), dtype=np.uint32), np.array(list(s.claim_post), dtype=np.uint32)

    @property
    def journal(self) -> bytes:
        r = self.ptr.contents
        return bytes(bytearray(r.journal[i] for i in range(int(r.journal_len))))

    def serialize(self) -> bytes:
        L = load_library()
        buf = C.POINTER(C.c_uint8)()
        n = C.c_size_t()
        _chk(L.b2z_receipt_serialize(self.ptr, C.byref(buf), C.byref(n)))
        out = C.string_at(buf, n.value)
        L.b2z_bytes_free(buf)
        return out

    @staticmethod
    def deserialize(data: bytes) -> "Receipt":
        p = C.POINTER(ReceiptStruct)()
        b = (C.c_uint8 * len(data)).from_buffer_copy(data) if data else (C.c_uint8 * 1)()
        _chk(load_library().b2z_receipt_deserialize(b, C.c_size_t(len(data)), C.byref(p)))
        return Receipt(p)

    @staticmethod
    def assemble(circuit: Circuit, seals, journal: bytes = b"") -> "Receipt":
        seals = [_u32(s) for s in seals]
        ptrs = (C.c_void_p * max(len(seals), 1))(*[s.ctypes.data for s in seals])
        lens = (C.c_size_t * max(len(seals), 1))(*[s.size for s in seals])
        p = C.POINTER(ReceiptStruct)()
        j = (C.c_uint8 * max(len(journal), 1)).from_buffer_copy(journal.ljust(1, b"\0"))
        _chk(load_library().b2z_receipt_assemble(circuit.h, ptrs, lens, C.c_size_t(len(seals)), j, C.c_size_t(len(journal)), C.byref(p)))
        return Receipt(p)

    def close(self):
        if self.ptr:
            load_library().b2z_receipt_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class _SessionCircuit(Circuit):
    """View of a session worker's circuit (owned by the session: never freed from here)."""

    def __init__(self, handle):
        L = load_library()
        self.h = C.c_void_p(handle)
        self.blob = None
        self.group_sizes = tuple(int(L.b2z_circuit_group_size(self.h, C.c_int(g))) for g in range(3))
        self.globals_size = int(L.b2z_circuit_globals_size(self.h))
        self.mix_size = int(L.b2z_circuit_mix_size(self.h))

    def close(self):
        self.h = C.c_void_p()


class _SessionHal(Hal):
    """View of a session worker's context (owned by the session)."""

    def __init__(self, L, ctx, device):
        self.L = L
        self.ctx = C.c_void_p(ctx)
        self.device = device

    def close(self):
        self.ctx = C.c_void_p()


class Session:
    """A prover server for one circuit on one GPU: `in_flight` segments proved concurrently (own streams), receipts in
    index order -- ProverServer::prove_session behind `default_prover().prove(env, elf)`."""

    def __init__(self, device: int, blob, in_flight: int = 2):
        L = load_library()
        self.L = L
        self.blob = _u32(blob)
        self.h = C.c_void_p()
        _chk(L.b2z_session_create(C.c_int(device), _vp(self.blob), C.c_size_t(self.blob.size), C.c_uint(in_flight), C.byref(self.h)))
        self.device = device
        self.in_flight = int(L.b2z_session_in_flight(self.h))
        self.circuit = _SessionCircuit(L.b2z_session_circuit(self.h))
        self.hals = [_SessionHal(L, L.b2z_session_context(self.h, C.c_uint(i)), device) for i in range(self.in_flight)]

    def set_control_root(self, po2: int, root):
        r = _u32(root)
        _chk(self.L.b2z_session_set_control_root(self.h, C.c_uint(po2), _vp(r)))

    def prove(self, segments, journal: bytes = b"") -> Receipt:
        """segments: list of dicts {po2, code, data, globals, accum_noise, zk_rows, on_device}; code/data/accum_noise are
        device pointers (ints) when on_device, else numpy arrays / host pointers."""
        keep = []

        def hp(x):
            if isinstance(x, np.ndarray):
                keep.append(x)
                return x.ctypes.data
            return int(x)
        arr = (SegmentInput * max(len(segments), 1))()
        for i, s in enumerate(segments):
            g = _u32(s["globals"])
            keep.append(g)
            arr[i] = SegmentInput(s.get("index", i), s["po2"], hp(s["code"]), hp(s["data"]), g.ctypes.data, hp(s["accum_noise"]),
                                  s["zk_rows"], 1 if s.get("on_device") else 0)
        p = C.POINTER(ReceiptStruct)()
        j = (C.c_uint8 * max(len(journal), 1)).from_buffer_copy(journal.ljust(1, b"\0"))
        _chk(self.L.b2z_prove_session(self.h, arr, C.c_size_t(len(segments)), j, C.c_size_t(len(journal)), C.byref(p)))
        return Receipt(p)

    def set_seal_cache(self, directory: str | None):
        """Seal cache / resume: proved seals are stored under `directory` and reused by later sessions."""
        _chk(self.L.b2z_session_set_seal_cache(self.h, directory.encode() if directory else None))

    def last_cache_hits(self) -> int:
        return int(self.L.b2z_session_last_cache_hits(self.h))

    def launch_count(self) -> int:
        return int(self.L.b2z_session_launch_count(self.h))

    def last_wall_ms(self) -> float:
        return float(self.L.b2z_session_last_wall_ms(self.h))

    def last_retries(self) -> int:
        return int(self.L.b2z_session_last_retries(self.h))

    def close(self):
        if self.h:
            self.L.b2z_session_destroy(self.h)
            self.h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
