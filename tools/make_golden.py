#!/usr/bin/env python3
"""Generate tests/golden/oracle_seals.json: SHA-256 of oracle seals on seeded synthetic witnesses.

These are REGRESSION vectors derived from the oracle itself (no seal of the real reference exists offline:
SURVEY.md §8c "parity unpinned"); they freeze the restated protocol so that later edits to the oracle or the
CUDA path cannot silently change the byte layout.  The GPU parity tests compare against the same digests."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import pyoracle as O  # noqa: E402
from zeth_b200 import synth_circuit as S  # noqa: E402

CASES = [("tiny", 5, 99), ("tiny", 9, 99), ("tiny", 10, 7), ("small", 11, 99), ("small", 13, 0xB2000002)]
MK = {"tiny": S.circuit_tiny, "small": S.circuit_small}
out = {"note": "derived from oracle/ (regression vectors, not reference golden vectors)", "cases": []}
for name, po2, seed in CASES:
    sc = MK[name]()
    oc = O.Circuit(sc.blob)
    n = 1 << po2
    zk = min(1994, n // 4)
    code, data, glob, noise = oc.witgen(po2, seed, zk)
    seal = oc.prove(po2, code, data, glob, noise, zk)
    assert oc.verify(seal) == 0
    out["cases"].append({"circuit": name, "po2": po2, "seed": seed, "zk_rows": zk, "seal_words": int(seal.size),
                         "sha256": hashlib.sha256(seal.tobytes()).hexdigest(),
                         "blob_sha256": hashlib.sha256(sc.blob.tobytes()).hexdigest()})
os.makedirs(os.path.join(ROOT, "tests", "golden"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "tests", "golden", "oracle_seals.json"), "w"), indent=1)
print(json.dumps(out, indent=1))
