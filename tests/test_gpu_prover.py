"""Whole-segment parity: the seal produced by the CUDA pipeline (through the C ABI) must be word-for-word the
oracle's seal on the same cached witness, must verify, and a tampered seal must be rejected."""
import numpy as np
import pytest

from zeth_b200 import hal as H
from zeth_b200 import synth_circuit as S

pytestmark = pytest.mark.gpu


def _witness(oracle, sc, po2, seed):
    oc = oracle.Circuit(sc.blob)
    n = 1 << po2
    zk = min(1994, n // 4)
    code, data, glob, noise = oc.witgen(po2, seed, zk)
    return oc, code, data, glob, noise, zk


@pytest.mark.parametrize("mk,po2", [(S.circuit_tiny, 6), (S.circuit_tiny, 9), (S.circuit_small, 12), (S.circuit_small, 14)])
def test_seal_parity_host_api(hal, oracle, mk, po2):
    sc = mk()
    oc, code, data, glob, noise, zk = _witness(oracle, sc, po2, 11 + po2)
    want = oc.prove(po2, code, data, glob, noise, zk)
    c = H.Circuit(hal, sc.blob)
    c.set_control_root(po2, oc.control_roots[po2])  # the ORACLE's control root: the product verifier must agree with it
    got = hal.prove_segment(c, po2, code, data, glob, noise, zk)
    assert got.size == want.size
    diff = np.nonzero(got != want)[0]
    assert diff.size == 0, f"first differing seal word {diff[:5]} of {want.size}"
    assert oc.verify(got) == 0
    assert c.verify(got) == 0  # product-side host verifier (receipt.verify)
    bad = got.copy()
    bad[bad.size // 3] ^= 1
    assert oc.verify(bad) != 0 and c.verify(bad) != 0
    c.close()


def test_witgen_accumulate_eval_check_parity(hal, oracle):
    sc = S.circuit_small()
    po2 = 11
    oc, code, data, glob, noise, zk = _witness(oracle, sc, po2, 77)
    n = 1 << po2
    wa, wc, wd = sc.group_sizes
    c = H.Circuit(hal, sc.blob)
    dcode, ddata = hal.alloc(wc * n), hal.alloc(wd * n)
    dglob, dnoise = hal.alloc(sc.globals_size), hal.alloc(wa * zk)
    hal.witgen(c, po2, dcode, ddata, dglob, dnoise, 77, zk)
    assert np.array_equal(hal.view(dcode, wc * n), code)
    assert np.array_equal(hal.view(ddata, wd * n), data)
    assert np.array_equal(hal.view(dglob, sc.globals_size), glob)
    assert np.array_equal(hal.view(dnoise, wa * zk), noise)
    rng = np.random.default_rng(5)
    mix = rng.integers(0, 15 * (1 << 27) + 1, size=sc.mix_size, dtype=np.uint64).astype(np.uint32)
    dmix = hal.copy_from(mix)
    daccum = hal.alloc(wa * n)
    hal.accumulate(c, po2, daccum, dcode, ddata, dmix, dnoise, zk)
    accum = oc.accumulate(po2, code, data, mix, noise, zk)
    assert np.array_equal(hal.view(daccum, wa * n), accum)
    # eval_check on LDEs of the three groups
    groups_h = [oracle.batch_expand_into_evaluate_ntt(oracle.zk_shift(oracle.batch_interpolate_ntt(g, w), w), w, 2)
                for g, w in ((accum, wa), (code, wc), (data, wd))]
    dgs = [hal.copy_from(g) for g in groups_h]
    pm = rng.integers(0, 15 * (1 << 27) + 1, size=4, dtype=np.uint64).astype(np.uint32)
    dcheck = hal.alloc(16 * n)
    hal.eval_check(c, dcheck, dgs, dglob, dmix, pm, po2)
    want = oc.eval_check(po2, groups_h, glob, mix, pm)
    assert np.array_equal(hal.view(dcheck, 16 * n), want)
    c.close()


def test_seal_parity_device_api_planning_circuit(hal, oracle):
    """BASELINE config-2 shape (W = 16 + 192 + 48) at a size the oracle proves in seconds."""
    sc = S.circuit_planning()
    po2 = 12
    oc, code, data, glob, noise, zk = _witness(oracle, sc, po2, 0xB2000002)
    want = oc.prove(po2, code, data, glob, noise, zk)
    c = H.Circuit(hal, sc.blob)
    dcode, ddata, dnoise = hal.copy_from(code), hal.copy_from(data), hal.copy_from(noise)
    got = hal.prove_segment_device(c, po2, dcode, ddata, glob, dnoise, zk)
    assert np.array_equal(got, want)
    assert oc.verify(got) == 0
    c.close()


def test_golden_seal_digests_on_gpu(hal, oracle):
    """The committed regression digests (tests/golden/oracle_seals.json) must be reproduced by the CUDA pipeline."""
    import hashlib
    import json
    import os
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_seals.json")))
    mk = {"tiny": S.circuit_tiny, "small": S.circuit_small}
    for case in gold["cases"]:
        sc = mk[case["circuit"]]()
        oc, code, data, glob, noise, zk = _witness(oracle, sc, case["po2"], case["seed"])
        c = H.Circuit(hal, sc.blob)
        seal = hal.prove_segment(c, case["po2"], code, data, glob, noise, zk)
        assert seal.size == case["seal_words"]
        assert hashlib.sha256(seal.tobytes()).hexdigest() == case["sha256"]
        c.close()


def test_full_gpu_pipeline_witgen_to_seal(hal, oracle):
    """Witness produced on the GPU (witgen kernels), proved on the GPU, checked by both verifiers; the oracle given the
    same seed must produce the same seal (end-to-end property at a size where the oracle still finishes quickly)."""
    sc = S.circuit_planning()
    po2, seed = 13, 0xB2000000 + 1000
    n = 1 << po2
    zk = min(1994, n // 4)
    wa, wc, wd = sc.group_sizes
    c = H.Circuit(hal, sc.blob)
    dcode, ddata = hal.alloc(wc * n), hal.alloc(wd * n)
    dglob, dnoise = hal.alloc(sc.globals_size), hal.alloc(wa * zk)
    hal.witgen(c, po2, dcode, ddata, dglob, dnoise, seed, zk)
    glob = hal.view(dglob, sc.globals_size)
    seal = hal.prove_segment_device(c, po2, dcode, ddata, glob, dnoise, zk)
    root = hal.compute_control_root(po2, dcode, wc)
    c.set_control_root(po2, root)
    assert c.verify(seal) == 0
    oc, code, data, oglob, noise, _ = _witness(oracle, sc, po2, seed)
    assert np.array_equal(root, oc.control_roots[po2])  # GPU-derived control root == the oracle's
    assert np.array_equal(glob, oglob)
    assert np.array_equal(seal, oc.prove(po2, code, data, oglob, noise, zk))
    c.close()


@pytest.mark.parametrize("po2", [13, 15, 16, 17, 18, 19, 20, 21, 22])
def test_baseline_size_segments_verify(hal, po2):
    """Every segment size between the protocol minimum (po2 = 13) and BASELINE.json's po2 = 20 / 22 (W = 256; each size
    exercises a different NTT pass plan and FRI round count): the oracle is too slow here, so use the size-independent
    property -- a seal produced from a GPU-generated valid witness must be accepted by the host verifier, must have the
    layout-determined length, and one flipped word must be rejected."""
    sc = S.circuit_planning()
    n = 1 << po2
    zk = 1994
    wa, wc, wd = sc.group_sizes
    c = H.Circuit(hal, sc.blob)
    dcode, ddata = hal.alloc(wc * n), hal.alloc(wd * n)
    dglob, dnoise = hal.alloc(sc.globals_size), hal.alloc(wa * zk)
    hal.witgen(c, po2, dcode, ddata, dglob, dnoise, 0xB2000002, zk)
    glob = hal.view(dglob, sc.globals_size)
    seal = hal.prove_segment_device(c, po2, dcode, ddata, glob, dnoise, zk)
    assert seal.size == c.seal_words(po2)
    assert c.verify(seal) == 15  # no control root yet: fails closed
    c.set_control_root(po2, hal.compute_control_root(po2, dcode, wc))
    assert c.verify(seal) == 0
    bad = seal.copy()
    bad[seal.size - 7] ^= 2
    assert c.verify(bad) != 0
    for p in (dcode, ddata, dglob, dnoise):
        hal.free(p)
    c.close()


def test_seal_parity_host_api_planning_circuit_chunked_leaves(hal, oracle):
    """Host-buffer API on the W = 256 circuit: the data group arrives in 32-column chunks and its leaves are hashed chunk by
    chunk (launch_hash_rows_chunk) -- the seal must still be the oracle's, word for word."""
    sc = S.circuit_planning()
    po2 = 12
    oc, code, data, glob, noise, zk = _witness(oracle, sc, po2, 0xB2000003)
    want = oc.prove(po2, code, data, glob, noise, zk)
    c = H.Circuit(hal, sc.blob)
    got = hal.prove_segment(c, po2, code, data, glob, noise, zk)
    assert np.array_equal(got, want)
    c.close()


def test_host_and_device_api_agree_at_mixed_kernel_sizes(hal):
    """po2 = 18 (2^20 leaves: the mixed INT / FP64 kernels run, with carried sponge state on the host path): the chunked
    host-buffer API and the one-shot device API must produce the same seal from the same witness."""
    sc = S.circuit_planning()
    po2, zk = 18, 1994
    n = 1 << po2
    wa, wc, wd = sc.group_sizes
    c = H.Circuit(hal, sc.blob)
    dcode, ddata = hal.alloc(wc * n), hal.alloc(wd * n)
    dglob, dnoise = hal.alloc(sc.globals_size), hal.alloc(wa * zk)
    hal.witgen(c, po2, dcode, ddata, dglob, dnoise, 0xB2000004, zk)
    glob = hal.view(dglob, sc.globals_size)
    code, data, noise = hal.view(dcode, wc * n), hal.view(ddata, wd * n), hal.view(dnoise, wa * zk)
    dev = hal.prove_segment_device(c, po2, dcode, ddata, glob, dnoise, zk)
    host = hal.prove_segment(c, po2, code, data, glob, noise, zk)
    assert np.array_equal(dev, host)
    c.set_control_root(po2, hal.compute_control_root(po2, dcode, wc))
    assert c.verify(host) == 0
    for p in (dcode, ddata, dglob, dnoise):
        hal.free(p)
    c.close()


# ---- BASELINE sizes: the shipping hot-path configuration (mixed INT/FP64 hash kernels, multi-pass NTT plans, chunked
# ---- leaf hashing) inside seals that are compared with the oracle ----
def _gpu_witness(hal, c, sc, po2, seed, zk=1994):
    n = 1 << po2
    wa, wc, wd = sc.group_sizes
    dcode, ddata = hal.alloc(wc * n), hal.alloc(wd * n)
    dglob, dnoise = hal.alloc(sc.globals_size), hal.alloc(wa * zk)
    hal.witgen(c, po2, dcode, ddata, dglob, dnoise, seed, zk)
    return dcode, ddata, dglob, dnoise, hal.view(dglob, sc.globals_size)


@pytest.mark.parametrize("po2,seed", [(18, 0xB2000012), (20, 0xB2000000 + 1000)])
def test_seal_parity_vs_live_oracle_at_baseline_sizes(hal, oracle, po2, seed):
    """po2 = 18 and po2 = 20 (BASELINE configs[2..3] segment size), W = 256: the whole seal, word for word, against the
    oracle run on the box's host cores -- through BOTH entry points (device-resident witness; host buffers, which are
    pageable numpy arrays here and therefore take the staged-upload path with chunk-carried leaf hashing)."""
    sc = S.circuit_planning()
    oc, code, data, glob, noise, zk = _witness(oracle, sc, po2, seed)
    want = oc.prove(po2, code, data, glob, noise, zk)
    assert oc.verify(want) == 0
    c = H.Circuit(hal, sc.blob)
    c.set_control_root(po2, oc.control_roots[po2])
    dcode, ddata, dnoise = hal.copy_from(code), hal.copy_from(data), hal.copy_from(noise)
    got = hal.prove_segment_device(c, po2, dcode, ddata, glob, dnoise, zk)
    diff = np.nonzero(got != want)[0]
    assert got.size == want.size and diff.size == 0, f"device API: first differing seal words {diff[:5]} of {want.size}"
    for p in (dcode, ddata, dnoise):
        hal.free(p)
    got = hal.prove_segment(c, po2, code, data, glob, noise, zk)
    diff = np.nonzero(got != want)[0]
    assert diff.size == 0, f"host API: first differing seal words {diff[:5]} of {want.size}"
    assert c.verify(got) == 0
    c.close()


def test_golden_seal_digests_at_baseline_sizes(hal):
    """BASELINE configs[1] ("Receipt bytes bit-exact vs CPU run"): the po2 = 22 seal of the planning circuit on seed
    0xB2000002 -- plus po2 18 / 20 -- must hash to the digest the ORACLE produced for the same cached witness
    (tests/golden/oracle_seals_big.json, frozen by tools/make_golden.py --big: the oracle needs ~10 minutes and ~35 GB
    for po2 = 22, so later runs only hash).  Device-resident and host-buffer (pageable -> staged) entry points."""
    import hashlib
    import json
    import os
    gold = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_seals_big.json")))
    sc = S.circuit_planning()
    assert {c["po2"] for c in gold["cases"]} >= {20, 22}
    for case in gold["cases"]:
        assert case["blob_sha256"] == hashlib.sha256(sc.blob.tobytes()).hexdigest()
        po2, zk = case["po2"], case["zk_rows"]
        n = 1 << po2
        wa, wc, wd = sc.group_sizes
        c = H.Circuit(hal, sc.blob)
        dcode, ddata, dglob, dnoise, glob = _gpu_witness(hal, c, sc, po2, case["seed"], zk)
        seal = hal.prove_segment_device(c, po2, dcode, ddata, glob, dnoise, zk)
        assert seal.size == case["seal_words"]
        assert hashlib.sha256(seal.tobytes()).hexdigest() == case["sha256"], f"po2 {po2}: device-API seal differs from the oracle's"
        c.set_control_root(po2, hal.compute_control_root(po2, dcode, wc))
        assert c.verify(seal) == 0
        code, data, noise = hal.view(dcode, wc * n), hal.view(ddata, wd * n), hal.view(dnoise, wa * zk)
        for p in (dcode, ddata, dglob, dnoise):
            hal.free(p)
        seal2 = hal.prove_segment(c, po2, code, data, glob, noise, zk)
        assert hashlib.sha256(seal2.tobytes()).hexdigest() == case["sha256"], f"po2 {po2}: host-API seal differs from the oracle's"
        del code, data, noise
        c.close()


def test_unsatisfied_witness_is_an_error_not_a_seal(hal, oracle):
    """prover.rs asserts exact DEEP divisions; a witness that violates a constraint must fail the call."""
    sc = S.circuit_small()
    po2 = 10
    oc, code, data, glob, noise, zk = _witness(oracle, sc, po2, 21)
    c = H.Circuit(hal, sc.blob)
    assert hal.prove_segment(c, po2, code, data, glob, noise, zk).size == c.seal_words(po2)
    derived = [col for col in range(sc.w_data) if sc.blob[-sc.w_data:][col] == 2]
    bad = data.copy()
    bad[derived[0] * (1 << po2) + 5] ^= 1
    with pytest.raises(H.B2ZError, match="does not satisfy"):
        hal.prove_segment(c, po2, code, bad, glob, noise, zk)
    dcode, ddata, dnoise = hal.copy_from(code), hal.copy_from(bad), hal.copy_from(noise)
    with pytest.raises(H.B2ZError, match="does not satisfy"):
        hal.prove_segment_device(c, po2, dcode, ddata, glob, dnoise, zk)
    # the context stays usable
    assert np.array_equal(hal.prove_segment(c, po2, code, data, glob, noise, zk), oc.prove(po2, code, data, glob, noise, zk))
    c.close()


def test_pinned_and_pageable_host_buffers_agree(hal, oracle):
    """b2z_prove_segment detects pageable sources and stages them through its pinned ring: same seal as from pinned memory."""
    import torch
    sc = S.circuit_planning()
    po2 = 14
    oc, code, data, glob, noise, zk = _witness(oracle, sc, po2, 0xB2000007)
    want = oc.prove(po2, code, data, glob, noise, zk)
    c = H.Circuit(hal, sc.blob)
    pageable = hal.prove_segment(c, po2, code, data, glob, noise, zk)
    pc, pd, pn = (torch.from_numpy(x.view(np.int32)).pin_memory() for x in (code, data, noise))
    pinned = hal.prove_segment(c, po2, pc.data_ptr(), pd.data_ptr(), glob, pn.data_ptr(), zk)
    assert np.array_equal(pageable, want) and np.array_equal(pinned, want)
    c.close()


def test_prove_session_receipts(hal, oracle):
    """ProverServer::prove_session behind the C ABI: 5 consecutive segments, two in flight, device-resident and host
    witnesses mixed; receipts come back in index order with the oracle's seals, the claim chain verifies, and a bad
    witness fails the session naming its segment."""
    sc = S.circuit_planning()
    po2, zk, seed0 = 12, 1024, 0xB2000000 + 1000
    oc = oracle.Circuit(sc.blob)
    sess = H.Session(0, sc.blob, in_flight=2)
    wit, want, keep = [], [], []
    for i in range(5):
        code, data, glob, noise = oc.witgen(po2, seed0 + i, zk)
        want.append(oc.prove(po2, code, data, glob, noise, zk))
        if i % 2 == 0:
            h = sess.hals[0]
            dc, dd, dn = h.copy_from(code), h.copy_from(data), h.copy_from(noise)
            keep += [dc, dd, dn]
            wit.append(dict(po2=po2, code=dc, data=dd, globals=glob, accum_noise=dn, zk_rows=zk, on_device=True))
        elif i == 1:  # pageable host memory: staged upload inside the call
            wit.append(dict(po2=po2, code=code, data=data, globals=glob, accum_noise=noise, zk_rows=zk, on_device=False))
        else:         # pinned host memory: the worker prefetches this witness while it proves its previous segment
            import torch
            pc, pd, pn = (torch.from_numpy(x.view(np.int32)).pin_memory() for x in (code, data, noise))
            keep += [pc, pd, pn]
            pinned_data3 = pd
            wit.append(dict(po2=po2, code=pc.data_ptr(), data=pd.data_ptr(), globals=glob, accum_noise=pn.data_ptr(), zk_rows=zk, on_device=False))
    sess.set_control_root(po2, oc.control_roots[po2])
    r = sess.prove(wit, journal=b"blockhash")
    assert r.n_segments == 5 and r.journal == b"blockhash"
    for i in range(5):
        assert r.segment(i).index == i
        assert np.array_equal(r.seal(i), want[i]), f"segment {i}"
    image_id = wit[0]["globals"][16:24]
    assert sess.circuit.verify_receipt(r, image_id) == (0, 0)
    r2 = H.Receipt.deserialize(r.serialize())
    assert sess.circuit.verify_receipt(r2, image_id) == (0, 0)
    # a violated constraint in segment 3 fails the whole session and names the segment
    derived = [col for col in range(sc.w_data) if sc.blob[-sc.w_data:][col] == 2]
    bad = pinned_data3.numpy().view(np.uint32).copy()  # segment 3's (pinned) data columns
    bad[derived[0] * (1 << po2) + 7] ^= 1
    wit_bad = list(wit)
    wit_bad[3] = dict(wit[3], data=bad)
    with pytest.raises(H.B2ZError, match="segment 3.*does not satisfy"):
        sess.prove(wit_bad)
    # and the session is still usable afterwards
    r3 = sess.prove(wit[:2])
    assert np.array_equal(r3.seal(1), want[1])
    sess.close()


# ---- circuits of the rv32im shape: one-hot mode selectors, AndCond-gated blocks, back taps up to 7 ----
@pytest.mark.parametrize("mk,po2", [(S.circuit_mux_small, 7), (S.circuit_mux_small, 11), (S.circuit_rv32im_scale, 10)])
def test_mux_circuit_seal_parity(hal, oracle, mk, po2):
    """GPU witgen == oracle witgen (one-hot column kinds), eval_check == oracle interpreter on the LDEs, whole seal ==
    oracle seal; for the rv32im-scale program (27 542 steps, 1 038 taps, 547 generated device functions) at po2 = 10."""
    sc = mk()
    oc, code, data, glob, noise, zk = _witness(oracle, sc, po2, 5 + po2)
    n = 1 << po2
    wa, wc, wd = sc.group_sizes
    c = H.Circuit(hal, sc.blob)
    dcode, ddata = hal.alloc(wc * n), hal.alloc(wd * n)
    dglob, dnoise = hal.alloc(sc.globals_size), hal.alloc(wa * zk)
    hal.witgen(c, po2, dcode, ddata, dglob, dnoise, 5 + po2, zk)
    assert np.array_equal(hal.view(dcode, wc * n), code)
    assert np.array_equal(hal.view(ddata, wd * n), data)
    assert np.array_equal(hal.view(dglob, sc.globals_size), glob)
    want = oc.prove(po2, code, data, glob, noise, zk)
    assert oc.verify(want) == 0
    got = hal.prove_segment_device(c, po2, dcode, ddata, glob, dnoise, zk)
    diff = np.nonzero(got != want)[0]
    assert diff.size == 0, f"first differing seal words {diff[:5]} of {want.size}"
    c.set_control_root(po2, oc.control_roots[po2])
    assert c.verify(got) == 0
    bad = data.copy()
    derived = [col for col in range(sc.w_data) if sc.blob[-sc.w_data:][col] == 2]
    bad[derived[1] * n + 3] ^= 1
    with pytest.raises(H.B2ZError, match="does not satisfy"):
        hal.prove_segment(c, po2, code, bad, glob, noise, zk)
    c.close()


def test_session_seal_cache_resumes(hal, oracle, tmp_path):
    """Seal cache / resume: the second run of a session finds every seal on disk (no kernel is launched), a
    corrupted cache file is detected by the verifier and only that segment is proved again."""
    sc = S.circuit_planning()
    po2, zk, seed0 = 11, 512, 0xB2000000 + 3000
    oc = oracle.Circuit(sc.blob)
    sess = H.Session(0, sc.blob, in_flight=2)
    wit, want = [], []
    for i in range(4):
        code, data, glob, noise = oc.witgen(po2, seed0 + i, zk)
        want.append(oc.prove(po2, code, data, glob, noise, zk))
        wit.append(dict(po2=po2, code=code, data=data, globals=glob, accum_noise=noise, zk_rows=zk, on_device=False))
    sess.set_control_root(po2, oc.control_roots[po2])
    sess.set_seal_cache(str(tmp_path))
    r1 = sess.prove(wit)
    assert sess.last_cache_hits() == 0 and len(list(tmp_path.glob("seal_*.bin"))) == 4
    launches = sess.launch_count()
    r2 = sess.prove(wit)
    assert sess.last_cache_hits() == 4 and sess.launch_count() == launches
    assert all(np.array_equal(r2.seal(i), want[i]) and np.array_equal(r1.seal(i), want[i]) for i in range(4))
    victim = sorted(tmp_path.glob("seal_*_2_*.bin"))[0]  # the file of segment index 2
    raw = bytearray(victim.read_bytes())
    raw[len(raw) // 2] ^= 1
    victim.write_bytes(bytes(raw))
    r3 = sess.prove(wit)
    assert sess.last_cache_hits() == 3 and sess.launch_count() > launches
    assert all(np.array_equal(r3.seal(i), want[i]) for i in range(4))
    sess.set_seal_cache(None)
    sess.prove(wit[:1])
    assert sess.last_cache_hits() == 0
    sess.close()
