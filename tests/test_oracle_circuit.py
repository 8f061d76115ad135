"""The oracle's two forms of CircuitHal::eval_check agree bit for bit: the one-point-at-a-time definition (every value an
ext element) and the lane-blocked, statically typed, slot-recycling form that the CPU baseline runs (oracle/circuit.c)."""
import numpy as np
import pytest

from oracle import pyoracle as O
from zeth_b200 import synth_circuit as S

P = 15 * (1 << 27) + 1


@pytest.mark.parametrize("mk,po2", [
    (S.circuit_tiny, 1), (S.circuit_tiny, 2), (S.circuit_tiny, 5), (S.circuit_small, 3), (S.circuit_small, 6),
    (S.circuit_mux_small, 4), (S.circuit_mux_small, 7), (S.circuit_planning, 5), (S.circuit_rv32im_scale, 4),
])
def test_lane_blocked_eval_check_equals_the_scalar_definition(mk, po2):
    O.build()
    sc = mk()
    oc = O.Circuit(sc.blob)
    rng = np.random.default_rng(0xB2 + po2)
    n = 1 << po2
    wa, wc, wd = sc.group_sizes
    # arbitrary (not constraint-satisfying) inputs: every op of the program sees generic values, incl. the wrap-around taps
    groups = [O.to_mont(rng.integers(0, P, size=max(w, 1) * 4 * n, dtype=np.uint64).astype(np.uint32)) for w in (wa, wc, wd)]
    glob = O.to_mont(rng.integers(0, P, size=max(sc.globals_size, 1), dtype=np.uint64).astype(np.uint32))
    mix = O.to_mont(rng.integers(0, P, size=max(sc.mix_size, 1), dtype=np.uint64).astype(np.uint32))
    poly_mix = O.to_mont(rng.integers(1, P, size=4, dtype=np.uint64).astype(np.uint32))
    a = oc.eval_check(po2, groups, glob, mix, poly_mix, scalar=True)
    b = oc.eval_check(po2, groups, glob, mix, poly_mix)
    assert a.shape == b.shape == (16 * n,)
    assert np.array_equal(a, b)
    assert (b < P).all()


def test_lane_blocked_hash_rows_and_hash_fold_equal_the_scalar_sponge():
    """16 sponges per permutation call (or_hash_rows / or_hash_fold) vs one at a time, incl. ragged row counts, empty rows and
    a partial last block"""
    import ctypes as C
    O.build()
    L = O.lib()
    rng = np.random.default_rng(5)
    vp = lambda a: a.ctypes.data_as(C.c_void_p)  # noqa: E731
    for rows, cols in ((1, 5), (7, 16), (16, 0), (33, 17), (64, 48), (1000, 192), (4096, 4)):
        m = rng.integers(0, P, size=max(rows * cols, 1), dtype=np.uint64).astype(np.uint32)
        a, b = np.zeros(rows * 8, dtype=np.uint32), np.zeros(rows * 8, dtype=np.uint32)
        L.or_hash_rows(vp(a), vp(m), C.c_size_t(rows), C.c_size_t(cols))
        L.or_hash_rows_scalar(vp(b), vp(m), C.c_size_t(rows), C.c_size_t(cols))
        assert np.array_equal(a, b), (rows, cols)
        if cols:  # and against the Python-level sponge of row 0
            assert np.array_equal(a[:8], O.hash_elem_slice(m.reshape(cols, rows)[:, 0]))
    for out in (1, 2, 8, 16, 17, 64, 100, 4096):
        nodes = rng.integers(0, P, size=8 * 4 * out, dtype=np.uint64).astype(np.uint32)
        n2 = nodes.copy()
        L.or_hash_fold(vp(nodes), C.c_size_t(2 * out), C.c_size_t(out))
        L.or_hash_fold_scalar(vp(n2), C.c_size_t(2 * out), C.c_size_t(out))
        assert np.array_equal(nodes, n2), out


def test_few_column_ntts_parallel_over_butterflies_equal_the_column_parallel_path():
    """count < host threads and n >= 4096 takes the per-layer parallel path (oracle/hal.c few_columns); the same columns inside
    a wide batch take the per-column path: identical results, and iNTT / LDE still invert each other"""
    O.build()
    nthreads = int(O.lib().or_num_threads())
    if nthreads < 2:
        pytest.skip("single host thread: only one path exists")
    rng = np.random.default_rng(11)
    n, few = 8192, 1
    wide = few + nthreads + 1
    x = rng.integers(0, P, size=wide * n, dtype=np.uint64).astype(np.uint32)
    a = x[: few * n].copy()
    O.batch_interpolate_ntt(a, few)
    b = x.copy()
    O.batch_interpolate_ntt(b, wide)
    assert np.array_equal(a, b[: few * n])
    ea = O.batch_expand_into_evaluate_ntt(a, few, 2)
    eb = O.batch_expand_into_evaluate_ntt(b, wide, 2)
    assert np.array_equal(ea, eb[: few * 4 * n])
    ra = a.copy()
    O.batch_evaluate_ntt(ra, few, 0)
    assert np.array_equal(ra, x[: few * n])
