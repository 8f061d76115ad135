// Interval-propagation PROOF of the magnitude envelope of the exact-FP64 Poseidon2 stream (zeth_b200/csrc/poseidon2_f64.cuh).
//
// The stream's digests equal the integer permutation's only if every floating-point operation in it is exact.  That
// rests on two lemmas and one global fact:
//   L1 (p2f_mm):  for integers a, b with |a b| / P + 1 < 2^49 and a_over_p within 1.5 * 2^-52 (relative) of a / P,
//                 the result is the integer a b - q P with |result| <= (0.5 + 1.5 * 2^-52 |a b| / P) P   [all 5 ops exact]
//   L2 (p2f_red): for an integer |x| < 2^52 the result is x - q P, |result| <= 0.51 P                      [both ops exact]
//   G:            every DADD / DFMA of the linear layers has integer operands and an integer result below 2^53.
// L1 / L2 are checked numerically on adversarial operands in tests/test_f64_stream.py.  THIS program proves G and the
// preconditions of L1 / L2 for EVERY call site of the real source: the header is compiled with `double` replaced by a type
// that carries an upper bound of |value| (the bounds do not depend on the data, only on the bounds of the inputs, so ONE
// abstract run covers all 2^744 states), and the two primitives are replaced by their contracts.
// Exit code 0 = every precondition holds; the report lists the tightest margins.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "../../zeth_b200/csrc/field.cuh"

static const long double TWO53 = 9007199254740992.0L, TWO52 = 4503599627370496.0L, TWO49 = 562949953421312.0L;
static const long double PP = 2013265921.0L;
static int g_fail = 0;
static long double g_max_mm_q = 0, g_max_lin = 0, g_max_red = 0;
static void require(bool ok, const char* what, long double v) {
  if (!ok) {
    fprintf(stderr, "VIOLATION: %s (value %.6Lg)\n", what, v);
    g_fail++;
  }
}

struct Iv {
  long double mag;  // upper bound of |value|
  bool integral;    // the value is an integer (all cell values and constants are); false: a quotient estimate x / P
  constexpr Iv() : mag(0), integral(true) {}
  constexpr Iv(long double m) : mag(m < 0 ? -m : m), integral(true) {}
  constexpr Iv(double m) : mag(m < 0 ? -(long double)m : (long double)m), integral((double)(long long)m == m) {}  // header constants / literals
  constexpr Iv(long double m, bool i) : mag(m < 0 ? -m : m), integral(i) {}
  Iv(uint32_t) : mag(PP - 1), integral(true) {}  // a canonical Montgomery word
  Iv operator-() const { return *this; }
  explicit operator int32_t() const {
    require(mag < 2147483648.0L, "double -> int32 conversion needs |x| < 2^31", mag);
    return 0;
  }
};
static Iv iv_add(Iv a, Iv b) {
  require(a.integral && b.integral, "DADD on a non-integer", 0);
  Iv r(a.mag + b.mag);
  require(r.mag < TWO53, "DADD result must stay below 2^53", r.mag);
  if (r.mag > g_max_lin) g_max_lin = r.mag;
  return r;
}
static Iv iv_mul(Iv a, Iv b) {
  if (!a.integral || !b.integral) return Iv(a.mag * b.mag, false);  // x * (1/P): a quotient estimate, consumed only by p2f_mm
  Iv r(a.mag * b.mag);
  require(r.mag < TWO53, "DMUL of integers must stay below 2^53", r.mag);
  return r;
}
static Iv iv_fma(Iv a, Iv b, Iv c) {
  require(a.integral && b.integral && c.integral, "DFMA outside the primitives must have integer operands", 0);
  Iv r(a.mag * b.mag + c.mag);
  require(r.mag < TWO53, "DFMA result must stay below 2^53", r.mag);
  if (r.mag > g_max_lin) g_max_lin = r.mag;
  return r;
}
#define B2Z_DADD(a, b) iv_add((a), (b))
#define B2Z_DMUL(a, b) iv_mul((a), (b))
#define B2Z_DFMA(a, b, c) iv_fma((a), (b), (c))
#define B2Z_P2F_PRIMS_HOOKED 1
// contracts of the two primitives (lemmas L1 / L2)
static Iv p2f_mm(Iv a, Iv a_over_p, Iv b) {
  require(a.integral && b.integral, "p2f_mm operands must be integers", 0);
  require(!a_over_p.integral || a_over_p.mag <= 1.0L, "p2f_mm: second argument must be the quotient estimate of the first", a_over_p.mag);
  const long double q = a.mag * b.mag / PP + 1;
  require(q < TWO49, "p2f_mm: |a b| / P + 1 < 2^49 (q (P - 1) exact)", q);
  if (q > g_max_mm_q) g_max_mm_q = q;
  return Iv((0.5L + 1.5L * a.mag * b.mag / PP / TWO52) * PP + 1);
}
static Iv p2f_red(Iv x) {
  require(x.integral, "p2f_red operand must be an integer", 0);
  require(x.mag < TWO52, "p2f_red: |x| < 2^52", x.mag);
  if (x.mag > g_max_red) g_max_red = x.mag;
  return Iv(0.51L * PP);
}

#define double Iv
#include "../../zeth_b200/csrc/poseidon2_f64.cuh"
#undef double

int main() {
  using namespace b2z;
  P2ConstsF64 k;
  for (int i = 0; i < P2_NUM_RC; i++) k.rc[i] = Iv(PP - 1);
  for (int i = 0; i < P2_CELLS; i++) {
    k.diag[i] = Iv(PP - 1);
    k.diag_over_p[i] = Iv(1.0L, false);
  }
  // a permutation starts from absorbed words (p2f_from_mont) and, for the carried capacity cells, p2f_red outputs
  Iv s[P2_CELLS];
  Iv worst_in(0.0L);
  for (int i = 0; i < P2_RATE; i++) {
    s[i] = p2f_from_mont(0u);
    if (s[i].mag > worst_in.mag) worst_in = s[i];
  }
  for (int i = P2_RATE; i < P2_CELLS; i++) s[i] = p2f_red(Iv(TWO52 - 1));
  p2f_permute(s, k);
  long double out = 0;
  for (int i = 0; i < P2_CELLS; i++)
    if (s[i].mag > out) out = s[i].mag;
  // squeeze / carry: conversion back to Montgomery words, and the per-block reduction of carried cells
  for (int i = 0; i < P2_CELLS; i++) (void)p2f_to_mont(s[i]);
  for (int i = P2_RATE; i < P2_CELLS; i++) (void)p2f_red(s[i]);
  printf("{\"violations\": %d, \"cell_in_bound_over_P\": %.4Lf, \"cell_out_bound_log2\": %.3Lf, \"max_linear_value_log2\": %.3Lf, "
         "\"max_p2f_mm_quotient_log2\": %.3Lf, \"max_p2f_red_input_log2\": %.3Lf}\n",
         g_fail, worst_in.mag / PP, log2l(out), log2l(g_max_lin), log2l(g_max_mm_q), log2l(g_max_red));
  return g_fail ? 1 : 0;
}
