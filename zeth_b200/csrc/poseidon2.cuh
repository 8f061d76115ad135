// Poseidon2 over BabyBear, width 24 (rate 16, output 8): the permutation shared by the device hash kernels
// (hash_rows / hash_fold) and the host-side Fiat-Shamir transcript.
//
// Algorithm = risc0-zkp 3.0.2 core/hash/poseidon2/mod.rs `poseidon2_mix`:
//   M_E ; 4 x {add rc[24]; x^7; M_E} ; 21 x {add rc to cell 0; x^7 on cell 0; M_I} ; 4 x {add rc; x^7; M_E}
// with constants from poseidon2_consts.h (generated + KAT-checked by tools/gen_poseidon2_consts.py).
// The state lives in 24 registers per thread; round constants sit in __constant__ memory in Montgomery form.
#pragma once
#include "field.cuh"

namespace b2z {

constexpr int P2_CELLS = 24;
constexpr int P2_RATE = 16;
constexpr int P2_OUT = 8;
constexpr int P2_ROUNDS_HALF_FULL = 4;
constexpr int P2_ROUNDS_PARTIAL = 21;
constexpr int P2_NUM_RC = 2 * P2_ROUNDS_HALF_FULL * P2_CELLS + P2_ROUNDS_PARTIAL;  // 213

struct P2Consts {
  uint32_t rc[P2_NUM_RC];          // Montgomery form
  uint32_t diag[P2_CELLS];         // Montgomery form
  int32_t rc_minus_p[P2_NUM_RC];   // rc - P as a signed value in (-P, 0]  (device s-box path)
  uint32_t diag_canon[P2_CELLS];   // canonical D_i                        (device Shoup path)
  uint32_t diag_shoup[P2_CELLS];   // floor(D_i * 2^32 / P)
};

const P2Consts& p2_host_consts();  // transcript.cpp

B2Z_HD uint32_t p2_sbox(uint32_t x) {
  uint32_t x2 = fp_mul(x, x);
  uint32_t x3 = fp_mul(x2, x);
  uint32_t x4 = fp_mul(x2, x2);
  return fp_mul(x3, x4);
}

// External linear layer: circ(2*M4, M4, ..., M4) with M4 = [[5,7,1,3],[4,6,1,1],[1,3,5,7],[1,1,4,6]].
B2Z_HD void p2_m_ext(uint32_t* s) {
#pragma unroll
  for (int c = 0; c < P2_CELLS; c += 4) {
    uint32_t a = s[c], b = s[c + 1], cc = s[c + 2], d = s[c + 3];
    uint32_t t0 = fp_add(a, b);
    uint32_t t1 = fp_add(cc, d);
    uint32_t t2 = fp_add(fp_add(b, b), t1);
    uint32_t t3 = fp_add(fp_add(d, d), t0);
    uint32_t t1x2 = fp_add(t1, t1);
    uint32_t t4 = fp_add(fp_add(t1x2, t1x2), t3);
    uint32_t t0x2 = fp_add(t0, t0);
    uint32_t t5 = fp_add(fp_add(t0x2, t0x2), t2);
    s[c] = fp_add(t3, t5);
    s[c + 1] = t5;
    s[c + 2] = fp_add(t2, t4);
    s[c + 3] = t4;
  }
  uint32_t sum0 = 0, sum1 = 0, sum2 = 0, sum3 = 0;
#pragma unroll
  for (int c = 0; c < P2_CELLS; c += 4) {
    sum0 = fp_add(sum0, s[c]);
    sum1 = fp_add(sum1, s[c + 1]);
    sum2 = fp_add(sum2, s[c + 2]);
    sum3 = fp_add(sum3, s[c + 3]);
  }
#pragma unroll
  for (int c = 0; c < P2_CELLS; c += 4) {
    s[c] = fp_add(s[c], sum0);
    s[c + 1] = fp_add(s[c + 1], sum1);
    s[c + 2] = fp_add(s[c + 2], sum2);
    s[c + 3] = fp_add(s[c + 3], sum3);
  }
}

// Internal linear layer: cell_i <- sum(cells) + D_i * cell_i.
B2Z_HD void p2_m_int(uint32_t* s, const uint32_t* diag) {
  uint32_t tot = 0;
#pragma unroll
  for (int i = 0; i < P2_CELLS; i++) tot = fp_add(tot, s[i]);
#pragma unroll
  for (int i = 0; i < P2_CELLS; i++) s[i] = fp_add(tot, fp_mul(diag[i], s[i]));
}

#if defined(__CUDACC__)
// ---- device-tuned variants (same canonical results; see DESIGN.md "Poseidon2 instruction budget") ----
// Pipe pinning.  The Montgomery products saturate the FMA pipe (IMAD.WIDE / IMAD.HI take ~2 issue slots each:
// tools/microbench/intpipe.cu), yet ptxas balances pipes by instruction COUNT and re-encodes about a quarter of the
// plain adds as IMAD.IADD on that same pipe.  `__viaddmin_u32(a, b, ones)` = min(a + b, 0xffffffff) = a + b is a single
// VIADDMNMX, which only exists on the ALU pipe; `ones` is a run-time all-ones kernel argument so it cannot be folded.
__device__ __forceinline__ uint32_t add_alu(uint32_t a, uint32_t b, uint32_t ones) { return __viaddmin_u32(a, b, ones); }
__device__ __forceinline__ uint32_t fp_add_p(uint32_t a, uint32_t b, uint32_t ones) {
  uint32_t r = add_alu(a, b, ones);
  return umin32(r, r - P);
}
// Signed-redundant Montgomery product: inputs in (-P, P) as int32, output in (-P, P); no final correction.
__device__ __forceinline__ int32_t fp_mul_s(int32_t a, int32_t b, uint32_t ones) {
  int64_t t = (int64_t)a * b;
  uint32_t m = (uint32_t)t * MONT_PINV;
  int32_t u = __mulhi((int32_t)m, (int32_t)P);
  return (int32_t)add_alu((uint32_t)(t >> 32), 0u - (uint32_t)u, ones);
}
// x^7 of (s + rc): one raw add (rc - P keeps the sum inside (-P, P)), four uncorrected products, one final correction
__device__ __forceinline__ uint32_t p2_sbox_rc(uint32_t s, int32_t rc_minus_p, uint32_t ones) {
  int32_t x = (int32_t)add_alu(s, (uint32_t)rc_minus_p, ones);
  int32_t x2 = fp_mul_s(x, x, ones);
  int32_t x3 = fp_mul_s(x2, x, ones);
  int32_t x4 = fp_mul_s(x2, x2, ones);
  uint32_t x7 = (uint32_t)fp_mul_s(x3, x4, ones);
  return umin32(x7, x7 + P);
}
// Shoup product by a constant: x * d mod P with dp = floor(d * 2^32 / P); x canonical, result canonical
__device__ __forceinline__ uint32_t fp_mul_shoup(uint32_t x, uint32_t d, uint32_t dp) {
  uint32_t q = __umulhi(x, dp);
  uint32_t r = x * d - q * P;
  return umin32(r, r - P);
}
__device__ __forceinline__ void p2_m_ext_dev(uint32_t* s, uint32_t ones) {
#pragma unroll
  for (int c = 0; c < P2_CELLS; c += 4) {
    uint32_t a = s[c], b = s[c + 1], cc = s[c + 2], d = s[c + 3];
    uint32_t t0 = fp_add_p(a, b, ones);
    uint32_t t1 = fp_add_p(cc, d, ones);
    uint32_t t2 = fp_add_p(fp_add_p(b, b, ones), t1, ones);
    uint32_t t3 = fp_add_p(fp_add_p(d, d, ones), t0, ones);
    uint32_t t1x2 = fp_add_p(t1, t1, ones);
    uint32_t t4 = fp_add_p(fp_add_p(t1x2, t1x2, ones), t3, ones);
    uint32_t t0x2 = fp_add_p(t0, t0, ones);
    uint32_t t5 = fp_add_p(fp_add_p(t0x2, t0x2, ones), t2, ones);
    s[c] = fp_add_p(t3, t5, ones);
    s[c + 1] = t5;
    s[c + 2] = fp_add_p(t2, t4, ones);
    s[c + 3] = t4;
  }
  uint32_t sum0 = s[0], sum1 = s[1], sum2 = s[2], sum3 = s[3];
#pragma unroll
  for (int c = 4; c < P2_CELLS; c += 4) {
    sum0 = fp_add_p(sum0, s[c], ones);
    sum1 = fp_add_p(sum1, s[c + 1], ones);
    sum2 = fp_add_p(sum2, s[c + 2], ones);
    sum3 = fp_add_p(sum3, s[c + 3], ones);
  }
#pragma unroll
  for (int c = 0; c < P2_CELLS; c += 4) {
    s[c] = fp_add_p(s[c], sum0, ones);
    s[c + 1] = fp_add_p(s[c + 1], sum1, ones);
    s[c + 2] = fp_add_p(s[c + 2], sum2, ones);
    s[c + 3] = fp_add_p(s[c + 3], sum3, ones);
  }
}
__device__ __forceinline__ void p2_m_int_dev(uint32_t* s, const P2Consts& k, uint32_t ones) {
  // pairwise tree (depth 5) instead of a 23-long dependent chain: same 23 adds, shorter critical path
  uint32_t t8[12];
#pragma unroll
  for (int i = 0; i < 12; i++) t8[i] = fp_add_p(s[2 * i], s[2 * i + 1], ones);
  uint32_t t4[6];
#pragma unroll
  for (int i = 0; i < 6; i++) t4[i] = fp_add_p(t8[2 * i], t8[2 * i + 1], ones);
  uint32_t t2[3];
#pragma unroll
  for (int i = 0; i < 3; i++) t2[i] = fp_add_p(t4[2 * i], t4[2 * i + 1], ones);
  uint32_t tot = fp_add_p(fp_add_p(t2[0], t2[1], ones), t2[2], ones);
#pragma unroll
  for (int i = 0; i < P2_CELLS; i++) s[i] = fp_add_p(tot, fp_mul_shoup(s[i], k.diag_canon[i], k.diag_shoup[i]), ones);
}
// `ones` must be 0xffffffff at run time (a kernel argument the compiler cannot fold)
__device__ __forceinline__ void p2_permute_dev(uint32_t* s, const P2Consts& k, uint32_t ones) {
  const int32_t* rc = k.rc_minus_p;
  p2_m_ext_dev(s, ones);
  for (int r = 0; r < P2_ROUNDS_HALF_FULL; r++) {
#pragma unroll
    for (int i = 0; i < P2_CELLS; i++) s[i] = p2_sbox_rc(s[i], rc[i], ones);
    rc += P2_CELLS;
    p2_m_ext_dev(s, ones);
  }
  for (int r = 0; r < P2_ROUNDS_PARTIAL; r++) {
    s[0] = p2_sbox_rc(s[0], rc[0], ones);
    rc += 1;
    p2_m_int_dev(s, k, ones);
  }
  for (int r = 0; r < P2_ROUNDS_HALF_FULL; r++) {
#pragma unroll
    for (int i = 0; i < P2_CELLS; i++) s[i] = p2_sbox_rc(s[i], rc[i], ones);
    rc += P2_CELLS;
    p2_m_ext_dev(s, ones);
  }
}
// Same permutation with ONE copy of each layer in the instruction stream (E; [21 x P after the 5th E]; S, nine times):
// ~14 KB of SASS instead of ~30 KB, so that it shares the SM's instruction cache with the FP64 stream
// (poseidon2_f64.cuh) in the mixed hash kernels.
__device__ __forceinline__ void p2_permute_dev_compact(uint32_t* s, const P2Consts& k, uint32_t ones) {
  const int32_t* rc = k.rc_minus_p;
#pragma unroll 1
  for (int e = 0; e < 2 * P2_ROUNDS_HALF_FULL + 1; e++) {
    p2_m_ext_dev(s, ones);
    if (e == 2 * P2_ROUNDS_HALF_FULL) break;
    if (e == P2_ROUNDS_HALF_FULL) {
#pragma unroll 1
      for (int r = 0; r < P2_ROUNDS_PARTIAL; r++) {
        s[0] = p2_sbox_rc(s[0], rc[r], ones);
        p2_m_int_dev(s, k, ones);
      }
      rc += P2_ROUNDS_PARTIAL;
    }
#pragma unroll
    for (int i = 0; i < P2_CELLS; i++) s[i] = p2_sbox_rc(s[i], rc[i], ones);
    rc += P2_CELLS;
  }
}
#endif

B2Z_HD void p2_permute(uint32_t* s, const P2Consts& k) {
  const uint32_t* rc = k.rc;
  p2_m_ext(s);
  for (int r = 0; r < P2_ROUNDS_HALF_FULL; r++) {
#pragma unroll
    for (int i = 0; i < P2_CELLS; i++) s[i] = p2_sbox(fp_add(s[i], rc[i]));
    rc += P2_CELLS;
    p2_m_ext(s);
  }
  for (int r = 0; r < P2_ROUNDS_PARTIAL; r++) {
    s[0] = p2_sbox(fp_add(s[0], rc[0]));
    rc += 1;
    p2_m_int(s, k.diag);
  }
  for (int r = 0; r < P2_ROUNDS_HALF_FULL; r++) {
#pragma unroll
    for (int i = 0; i < P2_CELLS; i++) s[i] = p2_sbox(fp_add(s[i], rc[i]));
    rc += P2_CELLS;
    p2_m_ext(s);
  }
}

}  // namespace b2z
