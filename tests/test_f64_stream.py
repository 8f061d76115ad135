"""Host-side properties of the exact-FP64 Poseidon2 stream (zeth_b200/csrc/poseidon2_f64.cuh), the second
instruction stream of the mixed hash kernels.  The header is plain C++ on the host (std::fma, no contraction), so its
arithmetic can be pinned against Python integers without a GPU: the GPU executes the same IEEE operations."""
import ctypes as C
import os
import random
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = 15 * (1 << 27) + 1
R = (1 << 32) % P

PROBE = r"""
#include "poseidon2_f64.cuh"
using namespace b2z;
extern "C" {
double probe_mm(double a, double b) { return p2f_mm(a, a * F64_INVP, b); }
double probe_sbox(double s, double rc) { return p2f_sbox_rc(s, rc); }
double probe_from_mont(unsigned x) { return p2f_from_mont(x); }
unsigned probe_to_mont(double x) { return p2f_to_mont(x); }
void probe_m_ext(double* s) { p2f_m_ext(s); }
double probe_red(double x) { return p2f_red(x); }
}
"""


@pytest.fixture(scope="module")
def probe(tmp_path_factory):
    d = tmp_path_factory.mktemp("f64probe")
    src, so = d / "probe.cpp", d / "probe.so"
    src.write_text(PROBE)
    subprocess.check_call(["/usr/bin/g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-mfma",
                           "-I", os.path.join(ROOT, "zeth_b200", "csrc"), "-o", str(so), str(src)])
    L = C.CDLL(str(so))
    L.probe_mm.restype = C.c_double
    L.probe_mm.argtypes = [C.c_double, C.c_double]
    L.probe_sbox.restype = C.c_double
    L.probe_sbox.argtypes = [C.c_double, C.c_double]
    L.probe_from_mont.restype = C.c_double
    L.probe_from_mont.argtypes = [C.c_uint]
    L.probe_to_mont.restype = C.c_uint
    L.probe_to_mont.argtypes = [C.c_double]
    L.probe_red.restype = C.c_double
    L.probe_red.argtypes = [C.c_double]
    return L


def test_modular_product_is_exact_over_its_whole_input_range(probe):
    """r = p2f_mm(a, a/P, b): an integer, congruent to a*b mod P, |r| <= 0.7 P, for |a| < 2^38, |b| < 2^40 (|ab| < 2^78)
    -- wider than anything the permutation feeds it (|cell| < 2^38 after a linear layer, constants < 2^31)."""
    rng = random.Random(64)
    edge = [0, 1, -1, P - 1, -(P - 1), P, (1 << 31) - 1, (1 << 38) - 1, -((1 << 38) - 1), (P - 1) // 2, (P + 1) // 2]
    cases = [(a, b) for a in edge for b in edge + [(1 << 40) - 1, -((1 << 40) - 1)]]
    for _ in range(20000):
        ea, eb = rng.randrange(1, 39), rng.randrange(1, 41)
        cases.append((rng.randrange(-(1 << ea), 1 << ea), rng.randrange(-(1 << eb), 1 << eb)))
    for a, b in cases:
        r = probe.probe_mm(float(a), float(b))
        assert r == int(r), (a, b, r)
        assert (int(r) - a * b) % P == 0, (a, b, r)
        assert abs(r) <= 0.7 * P, (a, b, r)


def test_sbox_and_montgomery_conversions(probe):
    rng = random.Random(7)
    for _ in range(5000):
        s = rng.randrange(-(1 << 38), 1 << 38)
        rc = rng.randrange(0, P)
        y = probe.probe_sbox(float(s), float(rc))
        assert y == int(y) and abs(y) <= 0.7 * P
        assert (int(y) - pow(s + rc, 7, P)) % P == 0
    for x in [0, 1, P - 1, R, 12345] + [rng.randrange(0, P) for _ in range(5000)]:
        plain = probe.probe_from_mont(x)  # x = plain * R mod P
        assert plain == int(plain) and (int(plain) * R - x) % P == 0
        assert probe.probe_to_mont(plain) == x  # canonical word back
    for v in [0, -1, 1, -(1 << 39), (1 << 39), P, -P, 5 * P + 3]:
        assert probe.probe_to_mont(float(v)) == (v * R) % P


def test_external_layer_matches_the_matrix(probe):
    """p2f_m_ext == circ(2 M4, M4, M4, M4, M4, M4) over the integers (no reduction inside), M4 as in SURVEY.md A.4."""
    m4 = np.array([[5, 7, 1, 3], [4, 6, 1, 1], [1, 3, 5, 7], [1, 1, 4, 6]], dtype=object)
    big = np.empty((24, 24), dtype=object)
    for i in range(6):
        for j in range(6):
            big[4 * i:4 * i + 4, 4 * j:4 * j + 4] = m4 * (2 if i == j else 1)
    rng = random.Random(3)
    for _ in range(200):
        s = [rng.randrange(-(1 << 31), 1 << 31) for _ in range(24)]
        buf = (C.c_double * 24)(*[float(v) for v in s])
        probe.probe_m_ext(buf)
        want = big.dot(np.array(s, dtype=object))
        assert [int(v) for v in buf] == list(want)


def test_interval_propagation_proves_the_envelope():
    """tests/cpp/f64_bounds.cpp compiles the REAL header with `double` replaced by a magnitude-tracking type and the two
    exactness-critical primitives replaced by their contracts: one abstract run visits every call site of the permutation,
    the Montgomery conversions and the carried-cell reduction, and proves (a) every linear-layer add / fma has integer
    operands and a result below 2^53, (b) every p2f_mm call satisfies |a b| / P + 1 < 2^49, (c) every p2f_red input is
    below 2^52.  The bounds depend only on the input bounds, so this covers all inputs -- not a sample."""
    import json
    import tempfile
    exe = os.path.join(tempfile.mkdtemp(), "f64_bounds")
    subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O1", "-o", exe, os.path.join(ROOT, "tests", "cpp", "f64_bounds.cpp")])
    out = subprocess.run([exe], capture_output=True, text=True)
    assert out.returncode == 0, out.stderr
    rep = json.loads(out.stdout)
    assert rep["violations"] == 0
    assert rep["cell_out_bound_log2"] < 38 and rep["max_linear_value_log2"] < 53 and rep["max_p2f_mm_quotient_log2"] < 49
    # the numeric lemma tests below must cover at least the operand range the proof found
    assert rep["max_p2f_mm_quotient_log2"] < 47


def test_primitive_contracts_on_adversarial_operands(probe):
    """Lemma L1 / L2 of tests/cpp/f64_bounds.cpp at and beyond the largest operands the proof found (quotients up to 2^47,
    i.e. |a b| up to 2^78), aimed at the rounding boundary: products within a few units of (q + 1/2) P."""
    rng = random.Random(99)
    n_hard = 0
    for _ in range(20000):
        ea = rng.randrange(20, 42)
        a = rng.randrange(1 << (ea - 1), 1 << ea) * rng.choice((1, -1))
        q = rng.randrange(1 << 30, 1 << 47)
        target = q * P + P // 2 + rng.randrange(-3, 4)  # a b ~ (q + 1/2) P: the worst case for the rounded quotient
        b = target // a + rng.randrange(-1, 2)
        if abs(b) >= 1 << 40 or abs(a * b) // P + 1 >= 1 << 49:
            continue
        n_hard += 1
        r = probe.probe_mm(float(a), float(b))
        assert r == int(r) and (int(r) - a * b) % P == 0 and abs(r) <= 0.75 * P, (a, b, r)
    assert n_hard > 3000
    for _ in range(20000):
        x = rng.randrange(-(1 << 52) + 1, 1 << 52)
        r = probe.probe_red(float(x))
        assert r == int(r) and (int(r) - x) % P == 0 and abs(r) <= 0.51 * P, (x, r)
    for q in [0, 1, -1, (1 << 21) - 1, -(1 << 21)]:
        for d in (-1, 0, 1):
            x = q * P + P // 2 + d
            r = probe.probe_red(float(x))
            assert (int(r) - x) % P == 0 and abs(r) <= 0.51 * P
