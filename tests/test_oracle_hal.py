"""Algebraic pins for the oracle's Hal ops (SURVEY.md §8c substitute oracle (ii), Appendix B.4):
naive DFT, round trips, expand == zero-padded transform, zk_shift == f(3x), FRI fold identity, Merkle re-hash."""
import numpy as np
import pytest

P = 15 * (1 << 27) + 1
ROU = {k: pow(137, 1 << (27 - k), P) for k in range(28)}


def brev(i, bits):
    return int(format(i, f"0{bits}b")[::-1], 2) if bits else 0


def rand(rng, n):
    return rng.integers(0, P, size=n, dtype=np.uint64)


def dft(coeffs, w):
    n = len(coeffs)
    return [sum(int(c) * pow(w, i * j, P) for j, c in enumerate(coeffs)) % P for i in range(n)]


@pytest.mark.parametrize("bits", [1, 3, 5])
def test_evaluate_ntt_is_dft_of_bitreversed_input(oracle, bits):
    rng = np.random.default_rng(bits)
    n = 1 << bits
    c = rand(rng, n)
    br = np.array([c[brev(i, bits)] for i in range(n)])
    got = oracle.from_mont(oracle.batch_evaluate_ntt(oracle.to_mont(br), 1, 0))
    assert list(got) == dft(c, ROU[bits])


@pytest.mark.parametrize("bits", [2, 4, 6])
def test_interpolate_is_inverse_dft_bitreversed_and_roundtrip(oracle, bits):
    rng = np.random.default_rng(10 + bits)
    n = 1 << bits
    c = rand(rng, 3 * n)
    ev = np.concatenate([dft(c[k * n:(k + 1) * n], ROU[bits]) for k in range(3)]).astype(np.uint64)
    got = oracle.from_mont(oracle.batch_interpolate_ntt(oracle.to_mont(ev), 3)).reshape(3, n)
    for k in range(3):
        assert list(got[k]) == [int(c[k * n + brev(i, bits)]) for i in range(n)]
    x = oracle.to_mont(rand(rng, 5 << 10))
    assert np.array_equal(oracle.batch_evaluate_ntt(oracle.batch_interpolate_ntt(x, 5), 5, 0), x)


def test_expand_equals_zero_padded_transform_and_zk_shift_is_3x(oracle):
    rng = np.random.default_rng(20)
    bits = 4
    n = 1 << bits
    c = rand(rng, n)
    br = oracle.to_mont(np.array([c[brev(i, bits)] for i in range(n)]))
    lde = oracle.from_mont(oracle.batch_expand_into_evaluate_ntt(br, 1, 2))
    w = ROU[bits + 2]
    assert list(lde) == [sum(int(cj) * pow(w, i * j, P) for j, cj in enumerate(c)) % P for i in range(4 * n)]
    shifted = oracle.from_mont(oracle.batch_expand_into_evaluate_ntt(oracle.zk_shift(br, 1), 1, 2))
    assert list(shifted) == [sum(int(cj) * pow(3 * pow(w, i, P), j, P) for j, cj in enumerate(c)) % P for i in range(4 * n)]


def test_bit_reverse_involution(oracle):
    rng = np.random.default_rng(30)
    x = rand(rng, 3 << 7).astype(np.uint32)
    y = oracle.batch_bit_reverse(x, 3)
    assert np.array_equal(oracle.batch_bit_reverse(y, 3), x)
    assert y[1] == x[64] and y[128 + 2] == x[128 + 32]


def ext_mul(a, b):
    r = [0] * 4
    for i in range(4):
        for j in range(4):
            k, v = i + j, a[i] * b[j]
            if k >= 4:
                k, v = k - 4, v * 11
            r[k] = (r[k] + v) % P
    return r


def test_fri_fold_matches_natural_fold(oracle):
    """out = bit-reversal of sum_i mix^i g_i, with f(x) = sum_i x^i g_i(x^16)."""
    rng = np.random.default_rng(40)
    count, bits = 8, 7
    n = 16 * count
    f = [[int(v) for v in rand(rng, 4)] for _ in range(n)]  # natural-order ext coefficients
    mix = [int(v) for v in rand(rng, 4)]
    g = [[f[16 * j + i] for j in range(count)] for i in range(16)]
    folded = [[0, 0, 0, 0] for _ in range(count)]
    mp = [1, 0, 0, 0]
    for i in range(16):
        for j in range(count):
            t = ext_mul(mp, g[i][j])
            folded[j] = [(folded[j][k] + t[k]) % P for k in range(4)]
        mp = ext_mul(mp, mix)
    inp = np.zeros(4 * n, dtype=np.uint64)
    for pos in range(n):
        for k in range(4):
            inp[k * n + pos] = f[brev(pos, bits)][k]
    out = oracle.from_mont(oracle.fri_fold(oracle.to_mont(inp), oracle.to_mont(np.array(mix)))).reshape(4, count)
    for pos in range(count):
        assert [int(out[k][pos]) for k in range(4)] == folded[brev(pos, 3)]


def test_batch_evaluate_any_is_horner(oracle):
    rng = np.random.default_rng(50)
    n, polys = 32, 3
    c = rand(rng, polys * n)
    which = np.array([2, 0, 1, 2], dtype=np.uint32)
    xs = [[int(v) for v in rand(rng, 4)] for _ in which]
    got = oracle.from_mont(oracle.batch_evaluate_any(oracle.to_mont(c), polys, which, oracle.to_mont(np.array(xs).reshape(-1)))).reshape(-1, 4)
    for e, (w, x) in enumerate(zip(which, xs)):
        tot = [0, 0, 0, 0]
        for j in reversed(range(n)):
            tot = ext_mul(tot, x)
            tot[0] = (tot[0] + int(c[w * n + j])) % P
        assert [int(v) for v in got[e]] == tot


def test_merkle_tree_paths_rehash(oracle):
    rng = np.random.default_rng(60)
    rows, cols = 64, 9
    m = oracle.to_mont(rand(rng, rows * cols))
    nodes = oracle.merkle_nodes(m, rows, cols).reshape(-1, 8)
    for idx in (0, 17, 63):
        cur = oracle.hash_elem_slice(m[idx::rows])
        node = idx + rows
        assert np.array_equal(cur, nodes[node])
        while node > 1:
            sib = nodes[node ^ 1]
            cur = oracle.hash_pair(sib, cur) if node & 1 else oracle.hash_pair(cur, sib)
            node //= 2
        assert np.array_equal(cur, nodes[1])


def test_ragged_and_empty_rows(oracle):
    z = oracle.hash_rows(np.zeros(0, dtype=np.uint32), 4, 0).reshape(4, 8)
    assert np.array_equal(z[0], oracle.hash_elem_slice(np.zeros(0, dtype=np.uint32))) and np.array_equal(z[0], z[3])
    rng = np.random.default_rng(61)
    for cols in (1, 15, 16, 17, 33):
        m = oracle.to_mont(rand(rng, 8 * cols))
        h = oracle.hash_rows(m, 8, cols).reshape(8, 8)
        assert np.array_equal(h[5], oracle.hash_elem_slice(m[5::8]))
