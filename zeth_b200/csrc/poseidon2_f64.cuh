// Poseidon2 over BabyBear computed EXACTLY in IEEE double arithmetic -- the second instruction stream of the mixed
// hash kernels (kernels_hash.cu).
//
// Why: the integer permutation (poseidon2.cuh) saturates the FMA-heavy (IMAD) and ALU pipes at ~72 % each while
// B200's FP64 pipe (64 DFMA / clk / SM, the same rate as IMAD) idles.  Warps that run THIS version of the same
// permutation draw on the FP64 pipe only, so INT warps and FP64 warps co-scheduled on one SM add their throughput.
// Same function, same bits out: every operation below is exact integer arithmetic carried in doubles
// (|values| < 2^53, products split with fma), so the digests are identical to p2_permute's.
//
// Representation: cells hold PLAIN residues (not Montgomery), integer-valued doubles, signed and unreduced:
//   * a modular product returns r with |r| <= 0.7 P (see p2f_mm),
//   * the linear layers add without any reduction (M_E grows a cell by at most x112, still < 2^38),
//   * M_I reduces the running total once per round so partial rounds stay bounded (|cell| < 1.3 P).
// Montgomery words are converted on absorb (x R^-1) and on squeeze (x R), 16 + 8 products per row block.
//
// Algorithm = risc0-zkp 3.0.2 core/hash/poseidon2/mod.rs poseidon2_mix, as poseidon2.cuh.
#pragma once
#include <cmath>

#include "poseidon2.cuh"

namespace b2z {

struct P2ConstsF64 {
  double rc[P2_NUM_RC];            // plain round constants in [0, P)
  double diag[P2_CELLS];           // plain D_i
  double diag_over_p[P2_CELLS];    // D_i / P (quotient estimate)
};

const P2ConstsF64& p2_host_consts_f64();  // transcript.cpp

constexpr double F64_P = 2013265921.0;
constexpr double F64_INVP = 1.0 / 2013265921.0;
constexpr double F64_MAGIC = 6755399441055744.0;    // 1.5 * 2^52: t + MAGIC - MAGIC == rint(t) for |t| < 2^51
constexpr double F64_NEG_PM1 = -2013265920.0;       // -(P - 1) = -15 * 2^27: q * (P - 1) is exact for q < 2^49
constexpr double F64_RINV = 943718400.0;            // R^-1 mod P, R = 2^32 (Montgomery word -> plain residue)
constexpr double F64_RINV_OVER_P = 943718400.0 / 2013265921.0;
constexpr double F64_RMOD = 268435454.0;            // R mod P (plain residue -> Montgomery word)
constexpr double F64_RMOD_OVER_P = 268435454.0 / 2013265921.0;

// (tests/cpp/f64_bounds.cpp re-includes this header with `double` replaced by a magnitude-tracking type and with the three
// macros and the two exactness-critical primitives p2f_mm / p2f_red hooked: it PROVES the envelope claims made below)
#if defined(B2Z_DADD)
// arithmetic supplied by the including translation unit
#elif defined(__CUDA_ARCH__)
#define B2Z_DADD(a, b) __dadd_rn((a), (b))
#define B2Z_DMUL(a, b) __dmul_rn((a), (b))
#define B2Z_DFMA(a, b, c) __fma_rn((a), (b), (c))
#else
#define B2Z_DADD(a, b) ((a) + (b))
#define B2Z_DMUL(a, b) ((a) * (b))
#define B2Z_DFMA(a, b, c) std::fma((a), (b), (c))
#endif

// a * b mod P with a_over_p ~= a / P (relative error <= 2^-51).  Needs |a * b| < 2^80.
//   q = rint(a b / P)                      (|q| < 2^49, off from the true nearest integer by < 0.2)
//   u = a b - q (P - 1)   exactly: one fma, the true value r + q is an integer below 2^50
//   r = u - q             in (-0.7 P, 0.7 P)
#if !defined(B2Z_P2F_PRIMS_HOOKED)
B2Z_HD double p2f_mm(double a, double a_over_p, double b) {
  double t = B2Z_DFMA(a_over_p, b, F64_MAGIC);
  double q = B2Z_DADD(t, -F64_MAGIC);
  double c = B2Z_DMUL(q, F64_NEG_PM1);
  double u = B2Z_DFMA(a, b, c);
  return B2Z_DADD(u, -q);
}
// x mod P for an integer |x| < 2^52: q = rint(x / P) has |q| < 2^22, so q * P is exact and one fma finishes; result in (-0.51 P, 0.51 P)
B2Z_HD double p2f_red(double x) {
  double q = B2Z_DADD(B2Z_DFMA(x, F64_INVP, F64_MAGIC), -F64_MAGIC);
  return B2Z_DFMA(q, -F64_P, x);
}
#endif

// (s + rc)^7: 24 FP64 operations
B2Z_HD double p2f_sbox_rc(double s, double rc) {
  double x = B2Z_DADD(s, rc);
  double xi = B2Z_DMUL(x, F64_INVP);
  double x2 = p2f_mm(x, xi, x);
  double x2i = B2Z_DMUL(x2, F64_INVP);
  double x3 = p2f_mm(x2, x2i, x);
  double x4 = p2f_mm(x2, x2i, x2);
  double x4i = B2Z_DMUL(x4, F64_INVP);
  return p2f_mm(x4, x4i, x3);
}

// External linear layer, no reductions: 8 operations per 4-cell group (2b + t1 etc. are single exact fmas).
B2Z_HD void p2f_m_ext(double* s) {
#pragma unroll
  for (int c = 0; c < P2_CELLS; c += 4) {
    double a = s[c], b = s[c + 1], cc = s[c + 2], d = s[c + 3];
    double t0 = B2Z_DADD(a, b);
    double t1 = B2Z_DADD(cc, d);
    double t2 = B2Z_DFMA(2.0, b, t1);
    double t3 = B2Z_DFMA(2.0, d, t0);
    double t4 = B2Z_DFMA(4.0, t1, t3);
    double t5 = B2Z_DFMA(4.0, t0, t2);
    s[c] = B2Z_DADD(t3, t5);
    s[c + 1] = t5;
    s[c + 2] = B2Z_DADD(t2, t4);
    s[c + 3] = t4;
  }
  double sum0 = s[0], sum1 = s[1], sum2 = s[2], sum3 = s[3];
#pragma unroll
  for (int c = 4; c < P2_CELLS; c += 4) {
    sum0 = B2Z_DADD(sum0, s[c]);
    sum1 = B2Z_DADD(sum1, s[c + 1]);
    sum2 = B2Z_DADD(sum2, s[c + 2]);
    sum3 = B2Z_DADD(sum3, s[c + 3]);
  }
#pragma unroll
  for (int c = 0; c < P2_CELLS; c += 4) {
    s[c] = B2Z_DADD(s[c], sum0);
    s[c + 1] = B2Z_DADD(s[c + 1], sum1);
    s[c + 2] = B2Z_DADD(s[c + 2], sum2);
    s[c + 3] = B2Z_DADD(s[c + 3], sum3);
  }
}

// Internal linear layer: cell_i <- (sum mod P) + D_i * cell_i mod P.  Cells stay below 1.3 P in magnitude.
B2Z_HD void p2f_m_int(double* s, const P2ConstsF64& k) {
  double t8[12];
#pragma unroll
  for (int i = 0; i < 12; i++) t8[i] = B2Z_DADD(s[2 * i], s[2 * i + 1]);
  double t4[6];
#pragma unroll
  for (int i = 0; i < 6; i++) t4[i] = B2Z_DADD(t8[2 * i], t8[2 * i + 1]);
  double tot = B2Z_DADD(B2Z_DADD(B2Z_DADD(t4[0], t4[1]), B2Z_DADD(t4[2], t4[3])), B2Z_DADD(t4[4], t4[5]));
  // tot mod P: |tot| < 2^37, so q * P is exact and one fma finishes the reduction
  double totr = p2f_red(tot);
#pragma unroll
  for (int i = 0; i < P2_CELLS; i++) s[i] = B2Z_DADD(p2f_mm(k.diag[i], k.diag_over_p[i], s[i]), totr);
}

// the permutation on plain-residue cells (|cell| < 2^31 on entry; |cell| < 2^38 on exit, reduce before exporting).
// One copy of each layer in the instruction stream: E; [21 x P after the 5th E]; S -- nine times (see
// p2_permute_dev_compact: both streams must fit the SM's instruction cache together).
B2Z_HD void p2f_permute(double* s, const P2ConstsF64& k) {
  const double* rc = k.rc;
#pragma unroll 1
  for (int e = 0; e < 2 * P2_ROUNDS_HALF_FULL + 1; e++) {
    p2f_m_ext(s);
    if (e == 2 * P2_ROUNDS_HALF_FULL) break;
    if (e == P2_ROUNDS_HALF_FULL) {
#pragma unroll 1
      for (int r = 0; r < P2_ROUNDS_PARTIAL; r++) {
        s[0] = p2f_sbox_rc(s[0], rc[r]);
        p2f_m_int(s, k);
      }
      rc += P2_ROUNDS_PARTIAL;
    }
#pragma unroll
    for (int i = 0; i < P2_CELLS; i++) s[i] = p2f_sbox_rc(s[i], rc[i]);
    rc += P2_CELLS;
  }
}

// canonical Montgomery word -> plain residue in (-0.7 P, 0.7 P)
B2Z_HD double p2f_from_mont(uint32_t xm) { return p2f_mm(F64_RINV, F64_RINV_OVER_P, (double)xm); }
// plain residue cell (|x| < 2^40) -> canonical Montgomery word
B2Z_HD uint32_t p2f_to_mont(double x) {
  double y = p2f_mm(F64_RMOD, F64_RMOD_OVER_P, x);  // in (-0.7 P, 0.7 P), an exact integer
  int32_t v = (int32_t)y;
  return v < 0 ? (uint32_t)(v + (int32_t)P) : (uint32_t)v;
}

}  // namespace b2z
