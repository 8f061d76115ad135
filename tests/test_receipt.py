"""Session-level containers on the CPU: SegmentReceipt / Receipt assembly, (de)serialisation and Receipt::verify
(what zeth runs at /root/reference crates/host/src/bin/cli.rs:103-107).  The seals come from the oracle (bit-identical to
the GPU prover's, tests/test_gpu_prover.py), so everything here runs without a GPU."""
import numpy as np
import pytest

from zeth_b200 import hal as H
from zeth_b200 import synth_circuit as S

PO2, ZK, SEED0 = 8, 64, 0xB2000000 + 1000


@pytest.fixture(scope="module")
def session(oracle):
    """Three consecutive segments of the planning circuit (32 globals: carries the pre/post claim words)."""
    sc = S.circuit_planning()
    oc = oracle.Circuit(sc.blob)
    seals, globs = [], []
    for i in range(3):
        code, data, glob, noise = oc.witgen(PO2, SEED0 + i, ZK)
        seals.append(oc.prove(PO2, code, data, glob, noise, ZK))
        globs.append(glob)
    c = H.Circuit(None, sc.blob)
    c.set_control_root(PO2, oc.control_roots[PO2])
    yield sc, oc, c, seals, globs
    c.close()


def test_receipt_assemble_verify_and_claim_chain(session):
    sc, oc, c, seals, globs = session
    r = H.Receipt.assemble(c, seals, journal=b"block-hash-0123456789abcdef")
    assert r.n_segments == 3 and r.journal == b"block-hash-0123456789abcdef"
    for i in range(3):
        s = r.segment(i)
        assert s.index == i and s.po2 == PO2 and s.hashfn == b"poseidon2"
        assert np.array_equal(r.seal(i), seals[i])
        pre, post = r.claim(i)
        assert np.array_equal(pre, globs[i][16:24]) and np.array_equal(post, globs[i][24:32])
        if i:
            assert np.array_equal(pre, r.claim(i - 1)[1])  # witgen: post(seed) == pre(seed + 1)
    image_id = globs[0][16:24]
    assert c.verify_receipt(r, image_id) == (0, 0)
    assert c.verify_receipt(r, None) == (0, 0)
    wrong = image_id.copy()
    wrong[0] ^= 1
    assert c.verify_receipt(r, wrong)[0] == 107


def test_receipt_serialisation_roundtrip(session):
    sc, oc, c, seals, globs = session
    r = H.Receipt.assemble(c, seals, journal=b"\x01\x02\x03")
    blob = r.serialize()
    r2 = H.Receipt.deserialize(blob)
    assert r2.serialize() == blob
    assert r2.n_segments == 3 and r2.journal == b"\x01\x02\x03"
    assert all(np.array_equal(r2.seal(i), seals[i]) for i in range(3))
    assert c.verify_receipt(r2, globs[0][16:24]) == (0, 0)
    for cut in (0, 3, 40, len(blob) // 2, len(blob) - 1):
        with pytest.raises(H.B2ZError):
            H.Receipt.deserialize(blob[:cut])
    with pytest.raises(H.B2ZError):
        H.Receipt.deserialize(blob + b"\x00")
    with pytest.raises(H.B2ZError):
        H.Receipt.deserialize(b"\xff" + blob[1:])


def test_receipt_verify_rejections(session):
    sc, oc, c, seals, globs = session
    # index order
    r = H.Receipt.assemble(c, seals)
    r.segment(1).index = 2
    assert c.verify_receipt(r) == (102, 1)
    # a tampered seal word -> the segment verifier's own code, and which segment
    bad = [s.copy() for s in seals]
    bad[2][bad[2].size // 2] ^= 4
    rc, seg = c.verify_receipt(H.Receipt.assemble(c, bad))
    assert 1 <= rc <= 13 and seg == 2
    # stored claim differs from what the seal's header says
    r = H.Receipt.assemble(c, seals)
    r.segment(0).claim_post[3] ^= 1
    assert c.verify_receipt(r) == (105, 0)
    # broken chain: segments 0 and 2 are not consecutive
    assert c.verify_receipt(H.Receipt.assemble(c, [seals[0], seals[2]])) == (106, 1)
    # wrong hash function name / verifier parameters
    r = H.Receipt.assemble(c, seals)
    r.segment(0).hashfn = b"sha-256"
    assert c.verify_receipt(r)[0] == 103
    r = H.Receipt.assemble(c, seals)
    r.ptr.contents.verifier_parameters[0] ^= 1
    assert c.verify_receipt(r)[0] == 104
    # empty receipt
    assert c.verify_receipt(H.Receipt.assemble(c, []))[0] == 101
    # a circuit without this po2's control root fails closed
    c2 = H.Circuit(None, sc.blob)
    assert c2.verify_receipt(H.Receipt.assemble(c2, seals))[0] == 15
    c2.close()
