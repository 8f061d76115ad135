"""Pin the oracle's field layer against the known-answer constants of SURVEY.md Appendix B.1
(risc0-core 3.0.0 field/baby_bear.rs; un-vendored, so these KATs are the anchor)."""
import numpy as np

P = 15 * (1 << 27) + 1
ROU_FWD = [1, 2013265920, 284861408, 1801542727, 567209306, 740045640, 918899846, 1881002012, 1453957774, 65325759, 1538055801,
           515192888, 483885487, 157393079, 1695124103, 2005211659, 1540072241, 88064245, 1542985445, 1269900459, 1461624142,
           825701067, 682402162, 1311873874, 1164520853, 352275361, 18769, 137]


def test_montgomery_constants(oracle):
    assert P * 0x88000001 % (1 << 32) == 1
    assert pow(2, 64, P) == 1172168163
    got = oracle.to_mont(np.array([1, 2, P - 1, 0, 11, 3]))
    assert [hex(int(x)) for x in got] == ["0xffffffe", "0x1ffffffc", "0x68000003", "0x0", "0x37ffffe9", "0x2ffffffa"]
    assert np.array_equal(oracle.from_mont(got), np.array([1, 2, P - 1, 0, 11, 3], dtype=np.uint32))


def test_rou_table_is_137_tower():
    for k, v in enumerate(ROU_FWD):
        assert v == pow(137, 1 << (27 - k), P)
    assert pow(137, 1 << 26, P) == P - 1  # order exactly 2^27
    assert pow(3, 1 << 27, P) != 1        # 3 is outside every 2-power subgroup


def test_mul_matches_python(oracle):
    rng = np.random.default_rng(1)
    a = rng.integers(0, P, size=4096, dtype=np.uint64)
    b = rng.integers(0, P, size=4096, dtype=np.uint64)
    got = oracle.from_mont(oracle.fp_mul(oracle.to_mont(a), oracle.to_mont(b)))
    assert np.array_equal(got, ((a * b) % P).astype(np.uint32))


def test_ext_field_x4_is_11_and_inverse(oracle):
    x = oracle.to_mont(np.array([0, 1, 0, 0]))
    x2 = oracle.fp4_mul(x, x)
    x4 = oracle.fp4_mul(x2, x2)
    assert list(oracle.from_mont(x4)) == [11, 0, 0, 0]
    rng = np.random.default_rng(2)
    a = oracle.to_mont(rng.integers(0, P, size=4 * 257, dtype=np.uint64))
    one = oracle.fp4_mul(a, oracle.fp4_inv(a)).reshape(-1, 4)
    assert np.array_equal(oracle.from_mont(one.reshape(-1)).reshape(-1, 4), np.tile(np.array([1, 0, 0, 0], dtype=np.uint32), (257, 1)))
