"""Op-by-op parity of the CUDA Hal (through the C ABI) against the CPU oracle -- the `dual` HAL pattern of
risc0-zkp hal/dual.rs: same seeded inputs, every output word must be identical."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
P = 15 * (1 << 27) + 1


def rand_fp(rng, n):
    return rng.integers(0, P, size=n, dtype=np.uint64).astype(np.uint32)


@pytest.mark.parametrize("po2,count", [(1, 3), (4, 5), (9, 7), (13, 3), (14, 2), (16, 3), (18, 2), (20, 1)])
def test_batch_interpolate_ntt(hal, oracle, po2, count):
    rng = np.random.default_rng(100 + po2)
    x = rand_fp(rng, count << po2)
    d = hal.copy_from(x)
    hal.batch_interpolate_ntt(d, count, po2)
    got = hal.view(d, x.size)
    hal.free(d)
    assert np.array_equal(got, oracle.batch_interpolate_ntt(x, count))


@pytest.mark.parametrize("po2,count", [(2, 3), (7, 5), (11, 7), (12, 2), (14, 3), (16, 2), (18, 1)])
def test_batch_expand_into_evaluate_ntt(hal, oracle, po2, count):
    rng = np.random.default_rng(200 + po2)
    x = rand_fp(rng, count << po2)
    d = hal.copy_from(x)
    o = hal.alloc(count << (po2 + 2))
    hal.batch_expand_into_evaluate_ntt(o, d, count, po2, 2)
    got = hal.view(o, count << (po2 + 2))
    hal.free(d)
    hal.free(o)
    assert np.array_equal(got, oracle.batch_expand_into_evaluate_ntt(x, count, 2))


@pytest.mark.parametrize("in_po2", [22, 24])
def test_expand_ntt_maximum_sizes(hal, oracle, in_po2):
    """LDE domains 2^24 (po2 = 22 segment) and 2^26 (po2 = 24, the largest the protocol allows): 3- and 4-pass plans."""
    rng = np.random.default_rng(250 + in_po2)
    x = rand_fp(rng, 1 << in_po2)
    d = hal.copy_from(x)
    o = hal.alloc(1 << (in_po2 + 2))
    hal.batch_expand_into_evaluate_ntt(o, d, 1, in_po2, 2)
    got = hal.view(o, 1 << (in_po2 + 2))
    want = oracle.batch_expand_into_evaluate_ntt(x, 1, 2)
    assert np.array_equal(got, want)
    # and back: interpolate the LDE, the first quarter (bit-reversed) holds the original coefficients
    hal.batch_interpolate_ntt(o, 1, in_po2 + 2)
    back = hal.view(o, 1 << (in_po2 + 2))
    hal.free(d)
    hal.free(o)
    assert np.array_equal(back, oracle.batch_interpolate_ntt(want, 1))


@pytest.mark.parametrize("po2,count", [(3, 2), (10, 3), (15, 2), (17, 1)])
def test_batch_evaluate_ntt_roundtrip(hal, oracle, po2, count):
    rng = np.random.default_rng(300 + po2)
    x = rand_fp(rng, count << po2)
    d = hal.copy_from(x)
    hal.batch_interpolate_ntt(d, count, po2)
    hal.batch_evaluate_ntt(d, count, po2, 0)
    got = hal.view(d, x.size)
    hal.free(d)
    assert np.array_equal(got, x)  # iNTT then NTT is the identity
    assert np.array_equal(oracle.batch_evaluate_ntt(oracle.batch_interpolate_ntt(x, count), count, 0), x)


@pytest.mark.parametrize("po2,count", [(1, 2), (8, 3), (13, 5), (16, 2)])
def test_bit_reverse_and_zk_shift(hal, oracle, po2, count):
    rng = np.random.default_rng(400 + po2)
    x = rand_fp(rng, count << po2)
    d = hal.copy_from(x)
    hal.batch_bit_reverse(d, count, po2)
    assert np.array_equal(hal.view(d, x.size), oracle.batch_bit_reverse(x, count))
    hal.batch_bit_reverse(d, count, po2)
    hal.zk_shift(d, count, po2)
    assert np.array_equal(hal.view(d, x.size), oracle.zk_shift(x, count))
    hal.free(d)


@pytest.mark.parametrize("rows,cols", [(16, 1), (64, 15), (256, 16), (512, 17), (1024, 64), (4096, 200), (8, 0)])
def test_hash_rows(hal, oracle, rows, cols):
    rng = np.random.default_rng(500 + rows + cols)
    m = rand_fp(rng, rows * cols)
    d = hal.copy_from(m if m.size else np.zeros(1, dtype=np.uint32))
    o = hal.alloc(rows * 8)
    hal.hash_rows(o, d, rows, cols)
    got = hal.view(o, rows * 8)
    hal.free(d)
    hal.free(o)
    assert np.array_equal(got, oracle.hash_rows(m, rows, cols))


@pytest.mark.parametrize("rows,cols", [(64, 5), (2048, 33)])
def test_merkle_build(hal, oracle, rows, cols):
    rng = np.random.default_rng(600 + rows)
    m = rand_fp(rng, rows * cols)
    d = hal.copy_from(m)
    nodes = hal.copy_from(np.zeros(2 * rows * 8, dtype=np.uint32))
    hal.merkle_build(nodes, d, rows, cols)
    got = hal.view(nodes, 2 * rows * 8)
    want = oracle.merkle_nodes(m, rows, cols)
    hal.free(d)
    hal.free(nodes)
    assert np.array_equal(got[8:], want[8:])  # node 0 is unused


@pytest.mark.parametrize("count", [1, 16, 1000, 4096])
def test_fri_fold(hal, oracle, count):
    rng = np.random.default_rng(700 + count)
    x = rand_fp(rng, 4 * 16 * count)
    mix = rand_fp(rng, 4)
    d = hal.copy_from(x)
    o = hal.alloc(4 * count)
    hal.fri_fold(o, d, mix, count)
    got = hal.view(o, 4 * count)
    hal.free(d)
    hal.free(o)
    assert np.array_equal(got, oracle.fri_fold(x, mix))


@pytest.mark.parametrize("po2,polys,evals", [(4, 3, 5), (12, 6, 9), (16, 4, 20)])
def test_batch_evaluate_any(hal, oracle, po2, polys, evals):
    rng = np.random.default_rng(800 + po2)
    coeffs = rand_fp(rng, polys << po2)
    which = rng.integers(0, polys, size=evals).astype(np.uint32)
    xs = rand_fp(rng, 4 * evals)
    dc, dw, dx = hal.copy_from(coeffs), hal.copy_from(which), hal.copy_from(xs)
    o = hal.alloc(4 * evals)
    hal.batch_evaluate_any(dc, polys, po2, dw, dx, evals, o)
    got = hal.view(o, 4 * evals)
    for p in (dc, dw, dx, o):
        hal.free(p)
    assert np.array_equal(got, oracle.batch_evaluate_any(coeffs, polys, which, xs))


@pytest.mark.parametrize("po2,cols,ncombos", [(5, 7, 3), (12, 40, 4), (14, 20, 11)])
def test_mix_poly_coeffs(hal, oracle, po2, cols, ncombos):
    rng = np.random.default_rng(900 + po2)
    n = 1 << po2
    inp = rand_fp(rng, cols * n)
    combos = rng.integers(0, ncombos, size=cols).astype(np.uint32)
    out0 = rand_fp(rng, 4 * n * ncombos)
    ms, mix = rand_fp(rng, 4), rand_fp(rng, 4)
    do, di, dcm = hal.copy_from(out0), hal.copy_from(inp), hal.copy_from(combos)
    hal.mix_poly_coeffs(do, ms, mix, di, dcm, cols, n)
    got = hal.view(do, out0.size)
    for p in (do, di, dcm):
        hal.free(p)
    assert np.array_equal(got, oracle.mix_poly_coeffs(out0, ms, mix, inp, combos, n))


def test_eltwise_and_gather(hal, oracle):
    rng = np.random.default_rng(1000)
    n = 5000
    a, b = rand_fp(rng, n), rand_fp(rng, n)
    da, db = hal.copy_from(a), hal.copy_from(b)
    o = hal.alloc(n)
    hal.eltwise_add_elem(o, da, db, n)
    assert np.array_equal(hal.view(o, n), ((a.astype(np.uint64) + b) % P).astype(np.uint32))
    ext = rand_fp(rng, 4 * 3 * 1000)
    de = hal.copy_from(ext)
    o2 = hal.alloc(4000)
    hal.eltwise_sum_extelem(o2, de, 1000, 3)
    assert np.array_equal(hal.view(o2, 4000), oracle.eltwise_sum_extelem(ext, 1000))
    z = a.copy()
    z[::7] = 0xFFFFFFFF
    dz = hal.copy_from(z)
    hal.eltwise_zeroize_elem(dz, n)
    assert np.array_equal(hal.view(dz, n), oracle.eltwise_zeroize_elem(z))
    g = hal.alloc(40)
    hal.gather_sample(g, da, 17, 40, 100)
    assert np.array_equal(hal.view(g, 40), a[17::100][:40])
    hal.eltwise_copy_elem(o, db, n)
    assert np.array_equal(hal.view(o, n), b)
    for p in (da, db, o, de, o2, dz, g):
        hal.free(p)


@pytest.mark.parametrize("count", [1, 7, 2048, 2049, 100000])
def test_poly_divide(hal, oracle, count):
    """q = p / (X - z): check q*(X - z) + r == p coefficient-wise with oracle ext arithmetic."""
    rng = np.random.default_rng(1100 + count)
    p = rand_fp(rng, 4 * count)
    z = rand_fp(rng, 4)
    d = hal.copy_from(p)
    rem = hal.poly_divide(d, count, z)
    q = hal.view(d, 4 * count).reshape(count, 4)
    hal.free(d)
    zq = oracle.fp4_mul(np.tile(z, count), q.reshape(-1)).reshape(count, 4).astype(np.int64)
    # p[i] = q[i-1] - z*q[i]  (q[-1] = 0, top quotient coefficient is 0) ; p[0] = r - z*q[0]
    shifted = np.vstack([rem.reshape(1, 4), q[:-1]]).astype(np.int64)
    recon = (shifted - zq) % P
    assert np.array_equal(recon.astype(np.uint32), p.reshape(count, 4))
    assert not q[-1].any()


def test_full_size_properties_po2_22(hal):
    """BASELINE full size (po2 = 22, LDE domain 2^24) without the oracle: linearity of the expand-NTT,
    interpolate(evaluate(x)) == x on the full LDE domain, and hash_rows depending only on its own row."""
    rng = np.random.default_rng(2200)
    po2 = 22
    n = 1 << po2
    a, b = rand_fp(rng, n), rand_fp(rng, n)
    s = ((a.astype(np.uint64) + b) % P).astype(np.uint32)
    da, db, ds = hal.copy_from(a), hal.copy_from(b), hal.copy_from(s)
    la, lb, ls = hal.alloc(4 * n), hal.alloc(4 * n), hal.alloc(4 * n)
    for src, dst in ((da, la), (db, lb), (ds, ls)):
        hal.batch_expand_into_evaluate_ntt(dst, src, 1, po2, 2)
    hal.eltwise_add_elem(la, la, lb, 4 * n)
    assert np.array_equal(hal.view(la, 4 * n), hal.view(ls, 4 * n))  # NTT(a) + NTT(b) == NTT(a + b)
    # round trip on the full LDE domain: interpolate(evaluate(x)) == x for bit-reversed-coefficient input x
    x = rand_fp(rng, 4 * n)
    dx = hal.copy_from(x)
    hal.batch_evaluate_ntt(dx, 1, po2 + 2, 0)
    hal.batch_interpolate_ntt(dx, 1, po2 + 2)
    assert np.array_equal(hal.view(dx, 4 * n), x)
    # hash_rows: digest of row r depends only on row r (compare a 2-column matrix with a row-permuted copy)
    rows = 1 << 16
    m = rand_fp(rng, 2 * rows)
    perm = rng.permutation(rows)
    mp = np.concatenate([m[:rows][perm], m[rows:][perm]])
    dm, dmp = hal.copy_from(m), hal.copy_from(mp)
    o1, o2 = hal.alloc(8 * rows), hal.alloc(8 * rows)
    hal.hash_rows(o1, dm, rows, 2)
    hal.hash_rows(o2, dmp, rows, 2)
    h1 = hal.view(o1, 8 * rows).reshape(rows, 8)
    h2 = hal.view(o2, 8 * rows).reshape(rows, 8)
    assert np.array_equal(h1[perm], h2)
    for p in (da, db, ds, la, lb, ls, dx, dm, dmp, o1, o2):
        hal.free(p)


@pytest.mark.gpu
@pytest.mark.parametrize("period,n_int", [(1, 0), (3, 2), (2, 1), (1, 1)])
@pytest.mark.parametrize("rows,cols", [(1 << 12, 40), (1 << 13, 16), (96, 5), (1 << 12, 0)])
def test_hash_rows_mixed_streams(hal, oracle, monkeypatch, rows, cols, period, n_int):
    """The mixed INT / FP64 leaf kernel (kernels_hash.cu): all-FP64, 2:1, 1:1 and all-INT role ratios give the oracle's digests."""
    monkeypatch.setenv("B2Z_HASH_MIXED_MIN_ROWS", "0")
    monkeypatch.setenv("B2Z_HASH_ROLE_PERIOD", str(period))
    monkeypatch.setenv("B2Z_HASH_ROLE_INT", str(n_int))
    rng = np.random.default_rng(rows * 31 + cols)
    m = rand_fp(rng, rows * cols)
    if cols:
        m[:rows] = np.where(np.arange(rows) % 3 == 0, P - 1, m[:rows])  # edge values in the first column
        m[rows - 1 :: rows] = 0
    dm = hal.copy_from(m) if cols else hal.alloc(1)
    do = hal.alloc(8 * rows)
    hal.hash_rows(do, dm, rows, cols)
    assert np.array_equal(hal.view(do, 8 * rows), oracle.hash_rows(m, rows, cols))
    hal.free(dm)
    hal.free(do)


@pytest.mark.gpu
def test_hash_rows_mixed_default_large(hal, oracle):
    """Default launch configuration at a size where the mixed kernel is the one that runs (2^20 rows)."""
    rows, cols = 1 << 20, 16
    m = rand_fp(np.random.default_rng(77), rows * cols)
    dm, do = hal.copy_from(m), hal.alloc(8 * rows)
    hal.hash_rows(do, dm, rows, cols)
    assert np.array_equal(hal.view(do, 8 * rows), oracle.hash_rows(m, rows, cols))
    hal.free(dm)
    hal.free(do)


@pytest.mark.parametrize("period,n_int", [(1, 0), (2, 1), (3, 1)])
def test_merkle_build_mixed_streams(hal, oracle, monkeypatch, period, n_int):
    """Leaves and interior layers through the mixed INT / FP64 kernels at every size (threshold forced to 0)."""
    monkeypatch.setenv("B2Z_HASH_MIXED_MIN_ROWS", "0")
    monkeypatch.setenv("B2Z_HASH_ROLE_PERIOD", str(period))
    monkeypatch.setenv("B2Z_HASH_ROLE_INT", str(n_int))
    rows, cols = 4096, 21
    m = rand_fp(np.random.default_rng(900 + period), rows * cols)
    d = hal.copy_from(m)
    nodes = hal.copy_from(np.zeros(2 * rows * 8, dtype=np.uint32))
    hal.merkle_build(nodes, d, rows, cols)
    got = hal.view(nodes, 2 * rows * 8)
    want = oracle.merkle_nodes(m, rows, cols)
    hal.free(d)
    hal.free(nodes)
    assert np.array_equal(got[8:], want[8:])


def test_merkle_build_mixed_default_large(hal, oracle):
    """2^20 leaves: the bottom interior layers run the mixed fold kernel in its default configuration."""
    rows, cols = 1 << 20, 3
    m = rand_fp(np.random.default_rng(901), rows * cols)
    d = hal.copy_from(m)
    nodes = hal.copy_from(np.zeros(2 * rows * 8, dtype=np.uint32))
    hal.merkle_build(nodes, d, rows, cols)
    got = hal.view(nodes, 2 * rows * 8)
    want = oracle.merkle_nodes(m, rows, cols)
    hal.free(d)
    hal.free(nodes)
    assert np.array_equal(got[8:], want[8:])


@pytest.mark.parametrize("mode", ["int_only", "mixed", "f64_only"])
@pytest.mark.parametrize("rows,chunks", [(1 << 12, (16, 32, 8)), (1 << 12, (32, 32)), (96, (16, 1)), (1 << 13, (48,))])
def test_hash_rows_chunk_carries_the_sponge(hal, oracle, monkeypatch, rows, chunks, mode):
    """b2z_hash_rows_chunk over consecutive column chunks == one hash_rows over all the columns, whichever stream resumes a row."""
    if mode == "int_only":
        monkeypatch.setenv("B2Z_HASH_MIXED", "0")
    else:
        monkeypatch.setenv("B2Z_HASH_MIXED_MIN_ROWS", "0")
        monkeypatch.setenv("B2Z_HASH_ROLE_PERIOD", "2" if mode == "mixed" else "1")
        monkeypatch.setenv("B2Z_HASH_ROLE_INT", "1" if mode == "mixed" else "0")
    cols = sum(chunks)
    m = rand_fp(np.random.default_rng(rows + cols), rows * cols)
    dm, do = hal.copy_from(m), hal.alloc(8 * rows)
    c0 = 0
    for k, c in enumerate(chunks):
        hal.hash_rows_chunk(do, dm + 4 * c0 * rows, rows, c, carry_in=k > 0, last=k == len(chunks) - 1)
        c0 += c
    assert np.array_equal(hal.view(do, 8 * rows), oracle.hash_rows(m, rows, cols))
    hal.free(dm)
    hal.free(do)


def test_hash_rows_chunk_rejects_ragged_middle_chunk(hal):
    do, dm = hal.alloc(8 * 64), hal.alloc(64 * 20)
    with pytest.raises(RuntimeError):
        hal.hash_rows_chunk(do, dm, 64, 20, carry_in=False, last=False)
    hal.free(do)
    hal.free(dm)


# ---- BASELINE config-5 shapes: 256-column launches (grid.y / 32-bit index guards), full compare against the oracle ----
def test_config5_shape_po2_20_cols_256(hal, oracle):
    """(po2, cols) = (20, 256): iNTT, x4 expand-NTT and hash_rows on the resulting [256][2^22] LDE matrix, every word
    against the oracle (1 GiB in, 4 GiB LDE, 2^22 leaves x 16 permutations)."""
    po2, cols = 20, 256
    n = 1 << po2
    rng = np.random.default_rng(0xB2000005)
    x = rng.integers(0, P, size=cols * n, dtype=np.uint32)  # uniform in [0, P) up to a 2^-32-scale bias: fine for parity
    d = hal.copy_from(x)
    hal.batch_interpolate_ntt(d, cols, po2)
    coeffs = hal.view(d, x.size)
    want_c = oracle.batch_interpolate_ntt(x, cols)
    assert np.array_equal(coeffs, want_c)
    del x, coeffs
    o = hal.alloc(cols * 4 * n)
    hal.batch_expand_into_evaluate_ntt(o, d, cols, po2, 2)
    want_lde = oracle.batch_expand_into_evaluate_ntt(want_c, cols, 2)
    del want_c
    # compare in column slabs to bound host memory
    slab = 32
    for c0 in range(0, cols, slab):
        got = np.empty(slab * 4 * n, dtype=np.uint32)
        import ctypes as C
        assert not hal.L.b2z_d2h(hal.ctx, got.ctypes.data_as(C.c_void_p), C.c_void_p(o + 4 * c0 * 4 * n), C.c_size_t(4 * got.size))
        assert np.array_equal(got, want_lde[c0 * 4 * n:(c0 + slab) * 4 * n]), f"LDE columns {c0}..{c0 + slab}"
    del got
    rows = 4 * n
    dig = hal.alloc(8 * rows)
    hal.hash_rows(dig, o, rows, cols)
    got_d = hal.view(dig, 8 * rows)
    assert np.array_equal(got_d, oracle.hash_rows(want_lde, rows, cols))
    for p in (d, o, dig):
        hal.free(p)


@pytest.mark.parametrize("out_size", [1, 2, 64, 512, 1024, 4096, 1 << 17, 1 << 20])
def test_hash_fold_direct(hal, oracle, out_size):
    """Hal::hash_fold(io, in_size, out_size) on its own: single-CTA sizes, thread-per-parent sizes and (2^20 parents) the
    mixed INT/FP64 persistent kernel."""
    rng = np.random.default_rng(900 + out_size.bit_length())
    nodes = np.zeros(8 * 4 * out_size, dtype=np.uint32)
    nodes[8 * 2 * out_size:] = rand_fp(rng, 8 * 2 * out_size)
    d = hal.copy_from(nodes)
    hal.hash_fold(d, 2 * out_size, out_size)
    got = hal.view(d, nodes.size)
    hal.free(d)
    want = oracle.hash_fold(nodes, 2 * out_size, out_size)
    assert np.array_equal(got, want)
    assert np.any(want[8 * out_size: 16 * out_size])


@pytest.mark.parametrize("n", [1, 2, 5, 1023, 1024, 1025, 70000, (1 << 20) + 3])
def test_prefix_products(hal, oracle, n):
    rng = np.random.default_rng(n)
    x = rand_fp(rng, 4 * n)
    d = hal.copy_from(x)
    hal.prefix_products(d, n)
    got = hal.view(d, 4 * n)
    hal.free(d)
    assert np.array_equal(got, oracle.prefix_products(x))


def test_scatter(hal, oracle):
    rng = np.random.default_rng(77)
    count, size = 5000, 1 << 16
    per = rng.integers(0, 6, size=count)
    index = np.concatenate([[0], np.cumsum(per)]).astype(np.uint32)
    total = int(index[-1])
    offsets = rng.permutation(size)[:total].astype(np.uint32)  # distinct targets: the result does not depend on thread order
    offsets[3] = offsets[2] if per[0] >= 4 else offsets[3]     # except inside one cycle, where the later entry wins
    values = rand_fp(rng, total)
    into = rand_fp(rng, size)
    di, dx, do_, dv = hal.copy_from(into), hal.copy_from(index), hal.copy_from(offsets), hal.copy_from(values)
    hal.scatter(di, dx, do_, dv, count)
    got = hal.view(di, size)
    for p in (di, dx, do_, dv):
        hal.free(p)
    assert np.array_equal(got, oracle.scatter(into, index, offsets, values))
    hal.scatter(0, 0, 0, 0, 0)  # empty: no-op


def test_negative_c_abi_arguments(hal):
    """Error returns (not crashes) for malformed calls through the C ABI."""
    import ctypes as C
    from zeth_b200 import hal as H
    L = hal.L
    d = hal.alloc(1024)
    assert b"exceeds" in L.b2z_batch_interpolate_ntt(hal.ctx, C.c_void_p(d), C.c_size_t(1), C.c_uint(27))
    assert b"distinct" in L.b2z_batch_expand_into_evaluate_ntt(hal.ctx, C.c_void_p(d), C.c_void_p(d), C.c_size_t(1), C.c_uint(4), C.c_uint(2))
    assert b"power of two" in L.b2z_merkle_build(hal.ctx, C.c_void_p(d), C.c_void_p(d), C.c_size_t(12), C.c_size_t(2))
    assert b"must not be empty" in L.b2z_hash_rows_chunk(hal.ctx, C.c_void_p(d), C.c_void_p(d), C.c_size_t(16), C.c_size_t(0), C.c_int(1), C.c_int(1))
    with pytest.raises(H.B2ZError, match="in_flight"):
        H.Session(0, np.zeros(64, dtype=np.uint32), in_flight=0)
    hal.free(d)


# ---- two-pass transforms (ntt_big.cuh): every size 2^20 .. 2^24, several columns, both directions ----
@pytest.mark.parametrize("in_po2,count", [(18, 5), (19, 3), (20, 3), (21, 2), (22, 2)])
def test_two_pass_expand_ntt(hal, oracle, in_po2, count):
    rng = np.random.default_rng(500 + in_po2)
    x = rand_fp(rng, count << in_po2)
    d = hal.copy_from(x)
    o = hal.alloc(count << (in_po2 + 2))
    hal.batch_expand_into_evaluate_ntt(o, d, count, in_po2, 2)
    got = hal.view(o, count << (in_po2 + 2))
    hal.free(d)
    hal.free(o)
    assert np.array_equal(got, oracle.batch_expand_into_evaluate_ntt(x, count, 2))


@pytest.mark.parametrize("po2,count", [(20, 5), (21, 3), (22, 3), (23, 2), (24, 1)])
def test_two_pass_interpolate_and_evaluate(hal, oracle, po2, count):
    rng = np.random.default_rng(600 + po2)
    x = rand_fp(rng, count << po2)
    d = hal.copy_from(x)
    hal.batch_interpolate_ntt(d, count, po2)
    got = hal.view(d, x.size)
    assert np.array_equal(got, oracle.batch_interpolate_ntt(x, count))
    hal.batch_evaluate_ntt(d, count, po2, 0)  # forward, in place, no expand: back to the evaluations
    assert np.array_equal(hal.view(d, x.size), x)
    hal.free(d)
