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
