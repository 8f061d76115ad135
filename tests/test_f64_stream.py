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
