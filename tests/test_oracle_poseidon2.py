"""Pin the oracle's Poseidon2 against the vectors of SURVEY.md Appendix B.2 / B.3
(risc0-zkp 3.0.2 core/hash/poseidon2; the 24-word KAT is the Horizen Labs t=24 BabyBear vector)."""
import hashlib
import os
import re
import struct

import numpy as np

P = 15 * (1 << 27) + 1
KAT_OUT = [0x2ed3e23d, 0x12921fb0, 0x0e659e79, 0x61d81dc9, 0x32bae33b, 0x62486ae3, 0x1e681b60, 0x24b91325, 0x2a2ef5b9, 0x50e8593e,
           0x5bc818ec, 0x10691997, 0x35a14520, 0x2ba6a3c5, 0x279d47ec, 0x55014e81, 0x5953a67f, 0x2f403111, 0x6b8828ff, 0x1801301f,
           0x2749207a, 0x3dc9cf21, 0x3c985ba2, 0x57a99864]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _consts(path):
    txt = open(path).read()
    body = txt[txt.index("P2_ROUND_CONSTANTS["):txt.index("P2_M_INT_DIAG")]
    return [int(x, 16) for x in re.findall(r"0x[0-9a-f]{8}", body)]


def test_round_constant_fingerprints_both_copies():
    for rel in ("oracle/poseidon2_consts.h", "zeth_b200/csrc/poseidon2_consts.h"):
        rc = _consts(os.path.join(ROOT, rel))
        assert len(rc) == 213
        assert rc[:4] == [0x0fa20c37, 0x0795bb97, 0x12c60b9c, 0x0eabd88e]
        assert rc[96:100] == [0x1da78ec2, 0x730b0924, 0x3eb56cf3, 0x5bd93073]
        assert sum(rc) % P == 1051808016
        assert hashlib.sha256(b"".join(struct.pack("<I", v) for v in rc)).hexdigest() == \
            "9f7aa102258e5f0e2fbfcb1a50b141bcb1914a08385a4c5640e1ae4dde1da983"


def test_permutation_kat(oracle):
    out = oracle.poseidon2_mix(oracle.to_mont(np.arange(24)))
    assert list(oracle.from_mont(out)) == KAT_OUT
    # first 8 as the raw Montgomery words that sit in digests
    assert [hex(int(x)) for x in out[:4]] == ["0x5419d919", "0x1061de89", "0x21495898", "0x4f4404bc"]


def test_derived_vectors(oracle):
    z = oracle.from_mont(oracle.poseidon2_mix(np.zeros(24, dtype=np.uint32)))[:8]
    assert list(z) == [0x39fa4dee, 0x386ee43e, 0x45e695ae, 0x24392948, 0x560b05e1, 0x009e435d, 0x4ae29966, 0x31a12e7b]
    s = oracle.to_mont(np.array(list(range(1, 17)) + [0] * 8))
    w = oracle.from_mont(oracle.poseidon2_mix(s))[:8]
    assert list(w) == [0x4680a52c, 0x2dabcb6a, 0x11971069, 0x0150fbad, 0x372c30e7, 0x1c0fc962, 0x23c11545, 0x1cfff986]


def test_sponge_framing(oracle):
    """unpadded_hash: exactly-one-block input == one permutation of [block || 0^8]; empty input == perm(0);
    a trailing partial block is zero padded (so [a] and [a, 0] collide by design, as upstream)."""
    blk = oracle.to_mont(np.arange(1, 17))
    want = oracle.poseidon2_mix(np.concatenate([blk, np.zeros(8, dtype=np.uint32)]))[:8]
    assert np.array_equal(oracle.hash_elem_slice(blk), want)
    assert np.array_equal(oracle.hash_elem_slice(np.zeros(0, dtype=np.uint32)), oracle.poseidon2_mix(np.zeros(24, dtype=np.uint32))[:8])
    a = oracle.to_mont(np.array([5]))
    assert np.array_equal(oracle.hash_elem_slice(a), oracle.hash_elem_slice(np.concatenate([a, [0]]).astype(np.uint32)))
    # two blocks: overwrite mode keeps the capacity cells of the first permutation
    two = oracle.to_mont(np.arange(1, 33))
    st = oracle.poseidon2_mix(np.concatenate([two[:16], np.zeros(8, dtype=np.uint32)]))
    st[:16] = two[16:]
    assert np.array_equal(oracle.hash_elem_slice(two), oracle.poseidon2_mix(st)[:8])


def test_hash_pair_is_one_block(oracle):
    rng = np.random.default_rng(3)
    a = rng.integers(0, P, size=8, dtype=np.uint64).astype(np.uint32)
    b = rng.integers(0, P, size=8, dtype=np.uint64).astype(np.uint32)
    assert np.array_equal(oracle.hash_pair(a, b), oracle.hash_elem_slice(np.concatenate([a, b])))


def test_rng(oracle):
    r1, r2 = oracle.Rng(), oracle.Rng()
    d = oracle.hash_elem_slice(oracle.to_mont(np.arange(3)))
    r1.mix(d)
    r2.mix(d)
    xs = [r1.random_elem() for _ in range(40)]
    assert xs == [r2.random_elem() for _ in range(40)]
    assert len(set(xs)) > 30 and all(x < P for x in xs)
    assert r1.random_bits(10) < 1024
    r3 = oracle.Rng()
    r3.mix(d)
    first = r3.random_elem()
    r3.mix(d)  # mixing while squeezing permutes first
    assert r3.random_elem() != first
