// BabyBear field (Montgomery form) and its quartic extension, shared by host and device code of the product.
//
// Same value encoding as risc0-core 3.0.0 field/baby_bear.rs (the un-vendored dependency zeth reaches through
// /root/reference crates/host/src/lib.rs:137): an Elem is the u32 x*2^32 mod P with P = 15*2^27+1, always
// reduced to [0, P); an ExtElem is 4 Elems over x^4 = 11.  The reduction sequence here is NOT upstream's
// (it is chosen for the sm_100a integer pipes: one IMAD.WIDE, one IMAD, one IMAD.HI and three ALU ops per
// product) but every function returns the unique canonical representative, so results are bit-identical.
#pragma once
#include <cstddef>
#include <cstdint>

#if defined(__CUDACC__)
#define B2Z_HD __host__ __device__ __forceinline__
#else
#define B2Z_HD inline
#endif

namespace b2z {

constexpr uint32_t P = 2013265921u;
constexpr uint32_t MONT_PINV = 0x88000001u;  // P^-1 mod 2^32
constexpr uint32_t MONT_R2 = 1172168163u;    // 2^64 mod P
constexpr uint32_t MONT_ONE = 0x0ffffffeu;   // 2^32 mod P
constexpr uint32_t MONT_BETA = 0x37ffffe9u;  // mont(11)
constexpr uint32_t MONT_THREE = 0x2ffffffau; // mont(3)
constexpr uint32_t INVALID = 0xffffffffu;

B2Z_HD uint32_t umin32(uint32_t a, uint32_t b) { return a < b ? a : b; }

B2Z_HD uint32_t fp_add(uint32_t a, uint32_t b) {
  uint32_t r = a + b;
  return umin32(r, r - P);
}
B2Z_HD uint32_t fp_sub(uint32_t a, uint32_t b) {
  uint32_t r = a - b;
  return umin32(r, r + P);
}
B2Z_HD uint32_t fp_neg(uint32_t a) { return fp_sub(0u, a); }

// Montgomery product a*b*2^-32 mod P.  t - m*P has a zero low word, so the result is hi(t) - hi(m*P) in (-P, P).
B2Z_HD uint32_t fp_mul(uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a * b;
  uint32_t m = (uint32_t)t * MONT_PINV;
#if defined(__CUDA_ARCH__)
  uint32_t u = __umulhi(m, P);
#else
  uint32_t u = (uint32_t)(((uint64_t)m * P) >> 32);
#endif
  uint32_t r = (uint32_t)(t >> 32) - u;
  return umin32(r, r + P);
}
B2Z_HD uint32_t fp_sqr(uint32_t a) { return fp_mul(a, a); }
// a0*b0 + a1*b1 + a2*b2 + a3*b3 (Montgomery products) with ONE reduction: every product of canonical words is < P^2 and
// 4 P^2 < 2^64, so the four terms accumulate in 64 bits (one IMAD.WIDE each); the high word (< 2^32 < 3 P) is brought below P
// by two conditional subtractions and one Montgomery step finishes: 10 instructions instead of 4 products + 3 adds (26).
B2Z_HD uint32_t fp_dot4(uint32_t a0, uint32_t b0, uint32_t a1, uint32_t b1, uint32_t a2, uint32_t b2, uint32_t a3, uint32_t b3) {
  uint64_t x = (uint64_t)a0 * b0 + (uint64_t)a1 * b1 + (uint64_t)a2 * b2 + (uint64_t)a3 * b3;
  uint32_t hi = (uint32_t)(x >> 32), lo = (uint32_t)x;
  hi = umin32(hi, hi - P);
  hi = umin32(hi, hi - P);
  uint32_t m = lo * MONT_PINV;
#if defined(__CUDA_ARCH__)
  uint32_t u = __umulhi(m, P);
#else
  uint32_t u = (uint32_t)(((uint64_t)m * P) >> 32);
#endif
  uint32_t r = hi - u;
  return umin32(r, r + P);
}
B2Z_HD uint32_t fp_from_u32(uint32_t x) { return fp_mul(x % P, MONT_R2); }
B2Z_HD uint32_t fp_to_u32(uint32_t a) { return fp_mul(a, 1u); }
B2Z_HD uint32_t fp_pow(uint32_t a, uint64_t n) {
  uint32_t r = MONT_ONE;
  while (n) {
    if (n & 1) r = fp_mul(r, a);
    a = fp_mul(a, a);
    n >>= 1;
  }
  return r;
}
B2Z_HD uint32_t fp_inv(uint32_t a) { return fp_pow(a, P - 2); }

struct Fp4 {
  uint32_t e[4];
};
B2Z_HD Fp4 fp4_make(uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  Fp4 r;
  r.e[0] = a; r.e[1] = b; r.e[2] = c; r.e[3] = d;
  return r;
}
B2Z_HD Fp4 fp4_zero() { return fp4_make(0, 0, 0, 0); }
B2Z_HD Fp4 fp4_one() { return fp4_make(MONT_ONE, 0, 0, 0); }
B2Z_HD Fp4 fp4_from_fp(uint32_t a) { return fp4_make(a, 0, 0, 0); }
B2Z_HD bool fp4_eq(const Fp4& a, const Fp4& b) { return a.e[0] == b.e[0] && a.e[1] == b.e[1] && a.e[2] == b.e[2] && a.e[3] == b.e[3]; }
B2Z_HD Fp4 fp4_add(const Fp4& a, const Fp4& b) {
  return fp4_make(fp_add(a.e[0], b.e[0]), fp_add(a.e[1], b.e[1]), fp_add(a.e[2], b.e[2]), fp_add(a.e[3], b.e[3]));
}
B2Z_HD Fp4 fp4_sub(const Fp4& a, const Fp4& b) {
  return fp4_make(fp_sub(a.e[0], b.e[0]), fp_sub(a.e[1], b.e[1]), fp_sub(a.e[2], b.e[2]), fp_sub(a.e[3], b.e[3]));
}
B2Z_HD Fp4 fp4_neg(const Fp4& a) { return fp4_make(fp_neg(a.e[0]), fp_neg(a.e[1]), fp_neg(a.e[2]), fp_neg(a.e[3])); }
B2Z_HD Fp4 fp4_add_fp(const Fp4& a, uint32_t b) { return fp4_make(fp_add(a.e[0], b), a.e[1], a.e[2], a.e[3]); }
B2Z_HD Fp4 fp4_sub_fp(const Fp4& a, uint32_t b) { return fp4_make(fp_sub(a.e[0], b), a.e[1], a.e[2], a.e[3]); }
B2Z_HD Fp4 fp4_mul_fp(const Fp4& a, uint32_t b) {
  return fp4_make(fp_mul(a.e[0], b), fp_mul(a.e[1], b), fp_mul(a.e[2], b), fp_mul(a.e[3], b));
}
// (a0 + a1 x + a2 x^2 + a3 x^3)(b0 + ...) mod x^4 - 11: every output component is a sum of four products -> one fp_dot4 each
// (3 products by beta + 4 single-reduction dot products instead of 19 products and 12 modular adds)
B2Z_HD Fp4 fp4_mul(const Fp4& a, const Fp4& b) {
  const uint32_t ba1 = fp_mul(MONT_BETA, a.e[1]), ba2 = fp_mul(MONT_BETA, a.e[2]), ba3 = fp_mul(MONT_BETA, a.e[3]);
  Fp4 r;
  r.e[0] = fp_dot4(a.e[0], b.e[0], ba1, b.e[3], ba2, b.e[2], ba3, b.e[1]);
  r.e[1] = fp_dot4(a.e[0], b.e[1], a.e[1], b.e[0], ba2, b.e[3], ba3, b.e[2]);
  r.e[2] = fp_dot4(a.e[0], b.e[2], a.e[1], b.e[1], a.e[2], b.e[0], ba3, b.e[3]);
  r.e[3] = fp_dot4(a.e[0], b.e[3], a.e[1], b.e[2], a.e[2], b.e[1], a.e[3], b.e[0]);
  return r;
}
B2Z_HD Fp4 fp4_pow(Fp4 a, uint64_t n) {
  Fp4 r = fp4_one();
  while (n) {
    if (n & 1) r = fp4_mul(r, a);
    a = fp4_mul(a, a);
    n >>= 1;
  }
  return r;
}
// Inverse through the quadratic tower F_p4 = F_p2[x]/(x^2 - y), y^2 = 11: a^-1 = conj(a) / (a * conj(a)).
B2Z_HD Fp4 fp4_inv(const Fp4& a) {
  uint32_t a2_0 = fp_add(fp_mul(a.e[0], a.e[0]), fp_mul(MONT_BETA, fp_mul(a.e[2], a.e[2])));
  uint32_t a2_1 = fp_mul(fp_add(a.e[0], a.e[0]), a.e[2]);
  uint32_t b2_0 = fp_add(fp_mul(a.e[1], a.e[1]), fp_mul(MONT_BETA, fp_mul(a.e[3], a.e[3])));
  uint32_t b2_1 = fp_mul(fp_add(a.e[1], a.e[1]), a.e[3]);
  uint32_t n0 = fp_sub(a2_0, fp_mul(MONT_BETA, b2_1));
  uint32_t n1 = fp_sub(a2_1, b2_0);
  uint32_t d = fp_sub(fp_mul(n0, n0), fp_mul(MONT_BETA, fp_mul(n1, n1)));
  uint32_t di = fp_inv(d);
  uint32_t i0 = fp_mul(n0, di);
  uint32_t i1 = fp_neg(fp_mul(n1, di));
  Fp4 conj = fp4_make(a.e[0], fp_neg(a.e[1]), a.e[2], fp_neg(a.e[3]));
  return fp4_mul(conj, fp4_make(i0, 0, i1, 0));
}

B2Z_HD unsigned ilog2(size_t n) {
  unsigned l = 0;
  while (((size_t)1 << l) < n) l++;
  return l;
}
B2Z_HD uint32_t bit_rev(uint32_t x, unsigned bits) {
#if defined(__CUDA_ARCH__)
  return bits ? (__brev(x) >> (32 - bits)) : 0u;
#else
  uint32_t r = 0;
  for (unsigned i = 0; i < bits; i++) r |= ((x >> i) & 1u) << (bits - 1 - i);
  return r;
#endif
}

// primitive 2^k-th roots of unity, canonical values (ROU_FWD[k] = 137^(2^(27-k)))
constexpr uint32_t ROU_FWD_CANON[28] = {1,          2013265920, 284861408,  1801542727, 567209306,  740045640,  918899846,
                                        1881002012, 1453957774, 65325759,   1538055801, 515192888,  483885487,  157393079,
                                        1695124103, 2005211659, 1540072241, 88064245,   1542985445, 1269900459, 1461624142,
                                        825701067,  682402162,  1311873874, 1164520853, 352275361,  18769,      137};
inline uint32_t rou_fwd(unsigned po2) { return fp_from_u32(ROU_FWD_CANON[po2]); }
inline uint32_t rou_rev(unsigned po2) { return fp_inv(rou_fwd(po2)); }

}  // namespace b2z
