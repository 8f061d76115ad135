"""Randomised (hypothesis) properties of the oracle's Hal ops: linearity, round trips, row independence."""
import numpy as np
from hypothesis import given, settings
from hypothesis import strategies as st

P = 15 * (1 << 27) + 1


def _rand(seed, n):
    return np.random.default_rng(seed).integers(0, P, size=n, dtype=np.uint64).astype(np.uint32)


def _add(a, b):
    return ((a.astype(np.uint64) + b) % P).astype(np.uint32)


@settings(max_examples=25, deadline=None)
@given(bits=st.integers(1, 12), count=st.integers(1, 5), seed=st.integers(0, 2**31))
def test_ntt_linearity_and_roundtrip(oracle, bits, count, seed):
    n = count << bits
    a, b = _rand(seed, n), _rand(seed + 1, n)
    ea, eb = oracle.batch_evaluate_ntt(a, count, 0), oracle.batch_evaluate_ntt(b, count, 0)
    assert np.array_equal(oracle.batch_evaluate_ntt(_add(a, b), count, 0), _add(ea, eb))
    assert np.array_equal(oracle.batch_interpolate_ntt(ea, count), a)
    lde = oracle.batch_expand_into_evaluate_ntt(a, count, 2)
    assert np.array_equal(oracle.batch_expand_into_evaluate_ntt(_add(a, b), count, 2), _add(lde, oracle.batch_expand_into_evaluate_ntt(b, count, 2)))
    # every 4th point of the x4 LDE (no shift) is the plain evaluation
    assert np.array_equal(lde.reshape(count, -1)[:, ::4].reshape(-1), ea)


@settings(max_examples=20, deadline=None)
@given(rows_log=st.integers(1, 8), cols=st.integers(0, 40), seed=st.integers(0, 2**31))
def test_hash_rows_row_independence(oracle, rows_log, cols, seed):
    rows = 1 << rows_log
    m = _rand(seed, rows * cols)
    perm = np.random.default_rng(seed + 7).permutation(rows)
    mp = m.reshape(cols, rows)[:, perm].reshape(-1) if cols else m
    h = oracle.hash_rows(m, rows, cols).reshape(rows, 8)
    hp = oracle.hash_rows(mp, rows, cols).reshape(rows, 8)
    assert np.array_equal(h[perm], hp)


@settings(max_examples=20, deadline=None)
@given(count_log=st.integers(0, 8), seed=st.integers(0, 2**31))
def test_fri_fold_linearity(oracle, count_log, seed):
    count = 1 << count_log
    a, b = _rand(seed, 64 * count), _rand(seed + 1, 64 * count)
    mix = _rand(seed + 2, 4)
    assert np.array_equal(oracle.fri_fold(_add(a, b), mix), _add(oracle.fri_fold(a, mix), oracle.fri_fold(b, mix)))
