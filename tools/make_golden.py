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
if "--big" not in sys.argv:
    json.dump(out, open(os.path.join(ROOT, "tests", "golden", "oracle_seals.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))
    sys.exit(0)

# --big [po2 ...]: BASELINE-size seals of the planning circuit (W = 256).  The oracle needs minutes and tens of GB for
# these, so their digests are frozen here once (run in the build container, 8 host threads) and the GPU tests only hash.
# po2 = 22 with seed 0xB2000002 is BASELINE.json configs[1] / configs[0]'s witness ("Receipt bytes bit-exact vs CPU run").
import time  # noqa: E402

BIG = {18: 0xB2000012, 20: 0xB2000000 + 1000, 22: 0xB2000002}
want = [int(a) for a in sys.argv[1:] if a.isdigit()] or sorted(BIG)
path = os.path.join(ROOT, "tests", "golden", "oracle_seals_big.json")
big = json.load(open(path)) if os.path.exists(path) else {
    "note": "derived from oracle/ on the planning circuit at BASELINE sizes (regression vectors; parity unpinned vs risc0)", "cases": []}
sc = S.circuit_planning()
oc = O.Circuit(sc.blob)
for po2 in want:
    seed = BIG[po2]
    t0 = time.time()
    code, data, glob, noise = oc.witgen(po2, seed, 1994)
    seal = oc.prove(po2, code, data, glob, noise, 1994)
    case = {"circuit": "planning", "po2": po2, "seed": seed, "zk_rows": 1994, "seal_words": int(seal.size),
            "sha256": hashlib.sha256(seal.tobytes()).hexdigest(), "blob_sha256": hashlib.sha256(sc.blob.tobytes()).hexdigest(),
            "oracle_seconds": round(time.time() - t0, 1), "oracle_threads": int(O.lib().or_num_threads())}
    del code, data, noise
    assert oc.verify(seal) == 0
    big["cases"] = [c for c in big["cases"] if c["po2"] != po2] + [case]
    big["cases"].sort(key=lambda c: c["po2"])
    json.dump(big, open(path, "w"), indent=1)
    print(json.dumps(case), flush=True)
