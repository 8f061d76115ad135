"""End-to-end pins for the oracle: prove -> verify accepts, tampering is rejected, the seal has the A.9 layout,
and committed golden digests (tests/golden/oracle_seals.json, made by tools/make_golden.py) stay stable."""
import hashlib
import json
import os

import numpy as np
import pytest

from zeth_b200 import synth_circuit as S

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_seals.json")


def _prove(oracle, mk, po2, seed):
    sc = mk()
    oc = oracle.Circuit(sc.blob)
    n = 1 << po2
    zk = min(1994, n // 4)
    code, data, glob, noise = oc.witgen(po2, seed, zk)
    return sc, oc, oc.prove(po2, code, data, glob, noise, zk), glob


@pytest.mark.parametrize("mk,po2", [(S.circuit_tiny, 5), (S.circuit_tiny, 8), (S.circuit_tiny, 9), (S.circuit_small, 11)])
def test_prove_verify_and_tamper(oracle, mk, po2):
    sc, oc, seal, glob = _prove(oracle, mk, po2, 99)
    assert oc.verify(seal) == 0
    # header = globals || po2 as raw Montgomery words
    assert np.array_equal(seal[: sc.globals_size], glob)
    assert oracle.from_mont(seal[sc.globals_size: sc.globals_size + 1])[0] == po2
    rng = np.random.default_rng(po2)
    for pos in list(rng.integers(0, seal.size, size=12)) + [0, seal.size - 1]:
        bad = seal.copy()
        bad[pos] ^= 1 << int(rng.integers(0, 27))
        assert oc.verify(bad) != 0, f"tampered word {pos} accepted"
    assert oc.verify(seal[:-1]) != 0
    assert oc.verify(np.concatenate([seal, [0]]).astype(np.uint32)) != 0


def test_invalid_witness_is_rejected(oracle):
    sc = S.circuit_tiny()
    oc = oracle.Circuit(sc.blob)
    po2, zk = 7, 32
    code, data, glob, noise = oc.witgen(po2, 5, zk)
    derived = [c for c in range(sc.w_data) if sc.blob[-sc.w_data:][c] == 2]
    data = data.copy()
    data[derived[0] * (1 << po2) + 3] ^= 1  # break one constrained cell on an active row
    seal = oc.prove(po2, code, data, glob, noise, zk)
    assert oc.verify(seal) != 0


def test_deterministic_and_seed_sensitive(oracle):
    _, _, a, _ = _prove(oracle, S.circuit_tiny, 7, 1)
    _, _, b, _ = _prove(oracle, S.circuit_tiny, 7, 1)
    _, _, c, _ = _prove(oracle, S.circuit_tiny, 7, 2)
    assert np.array_equal(a, b) and not np.array_equal(a, c)


def test_golden_seal_digests(oracle):
    gold = json.load(open(GOLDEN))
    mk = {"tiny": S.circuit_tiny, "small": S.circuit_small}
    for case in gold["cases"]:
        _, _, seal, _ = _prove(oracle, mk[case["circuit"]], case["po2"], case["seed"])
        assert seal.size == case["seal_words"]
        assert hashlib.sha256(seal.tobytes()).hexdigest() == case["sha256"]
