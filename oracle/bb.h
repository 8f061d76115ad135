/* ORACLE (test infrastructure, not product code): BabyBear field and its degree-4 extension.
 *
 * CPU restatement of risc0-core 3.0.0 `field/baby_bear.rs` (un-vendored crates.io dependency of
 * /root/reference, pinned at Cargo.lock:5338-5341; reached from crates/host/src/lib.rs:137).
 * Conventions (pinned by Appendix B.1):
 *   - P = 15*2^27 + 1; an `Elem` is a u32 holding x*2^32 mod P (Montgomery form), always < P;
 *   - `ExtElem` = 4 Elems, basis 1,x,x^2,x^3 with x^4 = 11 (BETA), NBETA = P - 11;
 *   - ROU_FWD[k] = 137^(2^(27-k)) is the primitive 2^k-th root of unity, ROU_REV[k] its inverse.
 * PARITY: field layer pinned by the known-answer constants of SURVEY.md B.1 (tests/test_oracle_field.py).
 */
#ifndef ORACLE_BB_H
#define ORACLE_BB_H
#include <stdint.h>
#include <stddef.h>

#define BB_P 2013265921u
#define BB_M 0x88000001u /* P^-1 mod 2^32 */
#define BB_R2 1172168163u /* 2^64 mod P */
#define BB_INVALID 0xffffffffu
#define BB_MAX_ROU_PO2 27

typedef uint32_t fp; /* Montgomery raw word */
typedef struct {
  fp e[4];
} fp4;

static inline fp fp_add(fp a, fp b) {
  uint32_t r = a + b;
  return r >= BB_P ? r - BB_P : r;
}
static inline fp fp_sub(fp a, fp b) {
  uint32_t r = a - b;
  return a < b ? r + BB_P : r;
}
static inline fp fp_neg(fp a) { return a ? BB_P - a : 0; }
/* Montgomery product: a*b*2^-32 mod P (baby_bear.rs `mul`). */
static inline fp fp_mul(fp a, fp b) {
  uint64_t o = (uint64_t)a * b;
  uint32_t low = 0u - (uint32_t)o;
  uint32_t red = BB_M * low;
  o += (uint64_t)red * BB_P;
  uint32_t r = (uint32_t)(o >> 32);
  return r >= BB_P ? r - BB_P : r;
}
/* canonical integer -> Montgomery word and back (Elem::new / as_u32). */
static inline fp fp_from_u32(uint32_t x) { return fp_mul(x % BB_P, BB_R2); }
static inline uint32_t fp_to_u32(fp a) { return fp_mul(a, 1); }
#define FP_ZERO 0u
#define FP_ONE 0x0ffffffeu /* mont(1) = 2^32 mod P */

static inline fp fp_pow(fp a, uint64_t n) {
  fp r = FP_ONE;
  while (n) {
    if (n & 1) r = fp_mul(r, a);
    a = fp_mul(a, a);
    n >>= 1;
  }
  return r;
}
static inline fp fp_inv(fp a) { return fp_pow(a, BB_P - 2); }

/* ---- extension field F_p[x]/(x^4 - 11) ---- */
static inline fp4 fp4_zero(void) {
  fp4 r = {{0, 0, 0, 0}};
  return r;
}
static inline fp4 fp4_one(void) {
  fp4 r = {{FP_ONE, 0, 0, 0}};
  return r;
}
static inline fp4 fp4_from_fp(fp a) {
  fp4 r = {{a, 0, 0, 0}};
  return r;
}
static inline int fp4_eq(fp4 a, fp4 b) {
  return a.e[0] == b.e[0] && a.e[1] == b.e[1] && a.e[2] == b.e[2] && a.e[3] == b.e[3];
}
static inline fp4 fp4_add(fp4 a, fp4 b) {
  fp4 r;
  for (int i = 0; i < 4; i++) r.e[i] = fp_add(a.e[i], b.e[i]);
  return r;
}
static inline fp4 fp4_sub(fp4 a, fp4 b) {
  fp4 r;
  for (int i = 0; i < 4; i++) r.e[i] = fp_sub(a.e[i], b.e[i]);
  return r;
}
static inline fp4 fp4_neg(fp4 a) {
  fp4 r;
  for (int i = 0; i < 4; i++) r.e[i] = fp_neg(a.e[i]);
  return r;
}
static inline fp4 fp4_mul_fp(fp4 a, fp b) {
  fp4 r;
  for (int i = 0; i < 4; i++) r.e[i] = fp_mul(a.e[i], b);
  return r;
}
#define FP_BETA 0x37ffffe9u /* mont(11) */
static inline fp4 fp4_mul(fp4 a, fp4 b) {
  const fp beta = FP_BETA;
  fp4 r;
  r.e[0] = fp_add(fp_mul(a.e[0], b.e[0]),
                  fp_mul(beta, fp_add(fp_add(fp_mul(a.e[1], b.e[3]), fp_mul(a.e[2], b.e[2])), fp_mul(a.e[3], b.e[1]))));
  r.e[1] = fp_add(fp_add(fp_mul(a.e[0], b.e[1]), fp_mul(a.e[1], b.e[0])),
                  fp_mul(beta, fp_add(fp_mul(a.e[2], b.e[3]), fp_mul(a.e[3], b.e[2]))));
  r.e[2] = fp_add(fp_add(fp_add(fp_mul(a.e[0], b.e[2]), fp_mul(a.e[1], b.e[1])), fp_mul(a.e[2], b.e[0])),
                  fp_mul(beta, fp_mul(a.e[3], b.e[3])));
  r.e[3] = fp_add(fp_add(fp_mul(a.e[0], b.e[3]), fp_mul(a.e[1], b.e[2])),
                  fp_add(fp_mul(a.e[2], b.e[1]), fp_mul(a.e[3], b.e[0])));
  return r;
}
static inline fp4 fp4_pow(fp4 a, uint64_t n) {
  fp4 r = fp4_one();
  while (n) {
    if (n & 1) r = fp4_mul(r, a);
    a = fp4_mul(a, a);
    n >>= 1;
  }
  return r;
}
/* Inverse via the tower F_p4 = F_p2[x]/(x^2 - y), y^2 = 11 (same construction as baby_bear.rs ExtElem::inv;
 * the inverse is unique so any correct method is bit-identical). */
static inline fp4 fp4_inv(fp4 a) {
  const fp beta = FP_BETA;
  /* a = A + x*B with A = a0 + a2 x^2, B = a1 + a3 x^2 (x^2 = y, y^2 = beta).  a * (A - xB) = A^2 - y B^2 in F_p2. */
  /* A^2 = (a0^2 + beta a2^2) + (2 a0 a2) y ;  B^2 = (a1^2 + beta a3^2) + (2 a1 a3) y ; y*B^2 = beta(2 a1 a3) + (a1^2 + beta a3^2) y */
  fp A2_0 = fp_add(fp_mul(a.e[0], a.e[0]), fp_mul(beta, fp_mul(a.e[2], a.e[2])));
  fp A2_1 = fp_mul(fp_add(a.e[0], a.e[0]), a.e[2]);
  fp B2_0 = fp_add(fp_mul(a.e[1], a.e[1]), fp_mul(beta, fp_mul(a.e[3], a.e[3])));
  fp B2_1 = fp_mul(fp_add(a.e[1], a.e[1]), a.e[3]);
  fp n0 = fp_sub(A2_0, fp_mul(beta, B2_1));
  fp n1 = fp_sub(A2_1, B2_0);
  /* norm to F_p: (n0 + n1 y)(n0 - n1 y) = n0^2 - beta n1^2 */
  fp d = fp_sub(fp_mul(n0, n0), fp_mul(beta, fp_mul(n1, n1)));
  fp di = fp_inv(d);
  /* inverse of (n0 + n1 y) = (n0 - n1 y) * di */
  fp i0 = fp_mul(n0, di);
  fp i1 = fp_neg(fp_mul(n1, di));
  /* result = (A - xB) * (i0 + i1 y), with A - xB = a0 - a1 x + a2 x^2 - a3 x^3 and y = x^2 */
  fp4 conj = {{a.e[0], fp_neg(a.e[1]), a.e[2], fp_neg(a.e[3])}};
  fp4 inv2 = {{i0, 0, i1, 0}};
  return fp4_mul(conj, inv2);
}

/* root-of-unity tables (canonical values from SURVEY.md B.1; converted on first use) */
extern const uint32_t BB_ROU_FWD_CANON[BB_MAX_ROU_PO2 + 1];
fp bb_rou_fwd(unsigned po2);
fp bb_rou_rev(unsigned po2);

static inline unsigned bb_log2(size_t n) {
  unsigned l = 0;
  while (((size_t)1 << l) < n) l++;
  return l;
}
static inline uint32_t bb_bit_rev(uint32_t x, unsigned bits) {
  uint32_t r = 0;
  for (unsigned i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
  return r;
}
#endif
