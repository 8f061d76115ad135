"""CPU-side checks of the product library: it loads, exports every symbol include/b2z.h declares, its host-side
Poseidon2 matches the KAT, circuits parse / generate / NVRTC-compile without a GPU, and compute fails loudly without one."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from zeth_b200 import hal as H
from zeth_b200 import synth_circuit as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared():
    txt = open(os.path.join(ROOT, "include", "b2z.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(b2z_[a-z0-9_]+)\s*\(", txt)))


def test_library_exports_every_declared_symbol():
    L = H.load_library()
    names = _declared()
    assert len(names) >= 40
    for n in names:
        assert hasattr(L, n), f"libb2z.so does not export {n}"
    assert b"sm_100a" in L.b2z_version()
    # the python mirror binds an error return for every error-returning entry point
    assert set(H._ERR_FUNCS) <= set(names)


def test_host_transcript_poseidon2_matches_kat():
    L = H.load_library()
    st = np.array([(i << 32) % (15 * (1 << 27) + 1) for i in range(24)], dtype=np.uint32)  # Montgomery form of 0..23
    L.b2z_host_poseidon2_mix(st.ctypes.data_as(C.c_void_p))
    assert [hex(int(x)) for x in st[:4]] == ["0x5419d919", "0x1061de89", "0x21495898", "0x4f4404bc"]


def test_host_hash_matches_oracle(oracle):
    L = H.load_library()
    rng = np.random.default_rng(1)
    for n in (0, 1, 16, 17, 100):
        e = rng.integers(0, 15 * (1 << 27) + 1, size=n, dtype=np.uint64).astype(np.uint32)
        d = np.zeros(8, dtype=np.uint32)
        L.b2z_host_hash_elem_slice(e.ctypes.data_as(C.c_void_p), C.c_size_t(n), d.ctypes.data_as(C.c_void_p))
        assert np.array_equal(d, oracle.hash_elem_slice(e))


def test_circuit_parse_codegen_sealsize_without_gpu(oracle):
    sc = S.circuit_small()
    c = H.Circuit(None, sc.blob)
    assert c.group_sizes == sc.group_sizes and c.mix_size == sc.mix_size
    src = c.source()
    assert "b2z_eval_check" in src and "b2z_witgen" in src
    oc = oracle.Circuit(sc.blob)
    for po2 in (6, 9, 13):
        n = 1 << po2
        zk = min(1994, n // 4)
        code, data, glob, noise = oc.witgen(po2, 3, zk)
        assert oc.prove(po2, code, data, glob, noise, zk).size == c.seal_words(po2)
    c.close()
    H.Circuit.precompile(sc.blob)  # NVRTC -> sm_100a cubin, no GPU needed


def test_bad_blob_is_rejected():
    sc = S.circuit_tiny()
    bad = sc.blob.copy()
    bad[0] ^= 1
    with pytest.raises(H.B2ZError):
        H.Circuit(None, bad)
    with pytest.raises(H.B2ZError):
        H.Circuit(None, sc.blob[:30])


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present; the loud-failure path is exercised on the CPU box")
    with pytest.raises(H.B2ZError, match="no usable CUDA device"):
        H.Hal(0)
    L = H.load_library()
    err = L.b2z_hash_rows(None, None, None, C.c_size_t(4), C.c_size_t(4))
    assert err and b"null context" in err


def test_product_verifier_matches_oracle_verifier(oracle):
    """b2z_verify_segment (product, host CPU) against the oracle's verifier: same accept / reject decisions and codes."""
    for mk, po2 in ((S.circuit_tiny, 6), (S.circuit_tiny, 10), (S.circuit_small, 12)):
        sc = mk()
        oc = oracle.Circuit(sc.blob)
        c = H.Circuit(None, sc.blob)
        n = 1 << po2
        zk = min(1994, n // 4)
        code, data, glob, noise = oc.witgen(po2, 31 + po2, zk)
        seal = oc.prove(po2, code, data, glob, noise, zk)
        assert c.verify(seal) == 15  # fails closed: no control root registered for this po2 yet
        c.set_control_root(po2, oc.control_roots[po2])
        assert c.verify(seal) == 0 and oc.verify(seal) == 0
        rng = np.random.default_rng(po2)
        for pos in list(rng.integers(0, seal.size, size=25)) + [0, sc.globals_size, seal.size - 1]:
            bad = seal.copy()
            bad[pos] ^= 1 << int(rng.integers(0, 31))
            assert c.verify(bad) == oc.verify(bad) != 0
        assert c.verify(seal[:-3]) == oc.verify(seal[:-3]) != 0
        c.close()


def test_host_poseidon2_f64_stream_is_bit_exact():
    """The exact-FP64 instruction stream of the mixed hash kernels (poseidon2_f64.cuh), host copy: same words as the
    integer permutation on random, all-zero, all-(P-1) and sparse states."""
    import ctypes as C

    import numpy as np
    from zeth_b200 import build

    L = C.CDLL(build.build())
    P = 15 * (1 << 27) + 1
    rng = np.random.default_rng(2024)
    for it in range(300):
        st = rng.integers(0, P, 24, dtype=np.uint64).astype(np.uint32)
        if it == 0:
            st[:] = 0
        elif it == 1:
            st[:] = P - 1
        elif it == 2:
            st[:] = 0
            st[5] = 1
        a, b = st.copy(), st.copy()
        L.b2z_host_poseidon2_mix(a.ctypes.data_as(C.c_void_p))
        L.b2z_host_poseidon2_mix_f64(b.ctypes.data_as(C.c_void_p))
        assert np.array_equal(a, b)


def test_verifier_binds_the_code_group_to_the_control_root(oracle):
    """receipt.verify(image_id) binds a seal to the PROGRAM (risc0-zkp verify/mod.rs check_code): a prover that commits
    to its own code columns -- e.g. the two selectors zeroed on every row, which makes every selector-gated
    constraint vanish -- must be rejected even though everything else in its seal is consistent."""
    sc = S.circuit_small()
    po2, zk = 9, 128
    oc = oracle.Circuit(sc.blob)
    c = H.Circuit(None, sc.blob)
    code, data, glob, noise = oc.witgen(po2, 17, zk)
    c.set_control_root(po2, oc.control_roots[po2])
    honest = oc.prove(po2, code, data, glob, noise, zk)
    assert c.verify(honest) == 0 == oc.verify(honest)
    forged_code = code.copy()
    forged_code[: 2 << po2] = 0  # code columns 0 / 1 = the `active` and `first` selectors -> 0 everywhere
    rng = np.random.default_rng(3)
    garbage = rng.integers(0, 15 * (1 << 27) + 1, size=data.size, dtype=np.uint64).astype(np.uint32)  # satisfies nothing
    forged = oc.prove(po2, forged_code, garbage, glob, noise, zk)
    assert c.verify(forged) == 14 == oc.verify(forged)
    # the same forged seal is "valid" for a verifier that trusts the prover's code columns: that is the hole being closed
    c2 = H.Circuit(None, sc.blob)
    c2.set_control_root(po2, oracle.control_root(po2, forged_code, sc.w_code))
    assert c2.verify(forged) == 0
    c.close()
    c2.close()


def test_noncanonical_digest_words_are_rejected(oracle):
    """Merkle digests are field elements: x + P hashes like x, so a verifier that accepts unreduced words lets several
    distinct seals verify for one proof (upstream's hash_pair asserts reduced inputs)."""
    P = 15 * (1 << 27) + 1
    sc = S.circuit_tiny()
    po2, zk = 8, 64
    oc = oracle.Circuit(sc.blob)
    c = H.Circuit(None, sc.blob)
    code, data, glob, noise = oc.witgen(po2, 5, zk)
    c.set_control_root(po2, oc.control_roots[po2])
    seal = oc.prove(po2, code, data, glob, noise, zk)
    assert c.verify(seal) == 0
    hits = 0
    top0 = sc.globals_size + 1  # first word of the code tree's top layer
    for pos in [top0, top0 + 9, seal.size - 3, seal.size - 20]:
        if int(seal[pos]) + P < (1 << 32):
            bad = seal.copy()
            bad[pos] = np.uint32(int(seal[pos]) + P)
            assert c.verify(bad) != 0 and oc.verify(bad) != 0
            hits += 1
    assert hits >= 2


def test_malformed_circuit_blobs_are_rejected():
    """circuit_parse validates every operand the verifier / generated kernels index with."""
    sc = S.circuit_tiny()
    blob = sc.blob
    n_taps, n_steps = int(blob[7]), int(blob[8])
    steps0 = 24 + n_taps
    ops = blob[steps0: steps0 + 5 * n_steps].reshape(-1, 5)
    gg = [i for i in range(n_steps) if ops[i, 0] == S.OP_GET_GLOBAL]
    assert gg
    for field, value in ((1, 2), (2, 1 << 20)):  # selector out of {globals, mix}; offset past the table
        bad = blob.copy()
        bad[steps0 + 5 * gg[0] + field] = value
        with pytest.raises(H.B2ZError, match="GetGlobal"):
            H.Circuit(None, bad)
    # duplicate back inside a register (taps stay sorted): DEEP interpolation would divide by zero
    taps = blob[24: 24 + n_taps]
    same = [i for i in range(1, n_taps) if (taps[i] & 0xff00ffff) == (taps[i - 1] & 0xff00ffff)]
    assert same
    bad = blob.copy()
    bad[24 + same[0]] = taps[same[0] - 1]
    with pytest.raises(H.B2ZError, match="strictly increasing"):
        H.Circuit(None, bad)
