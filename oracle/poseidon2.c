/* ORACLE (test infrastructure, not product code): Poseidon2 over BabyBear, width 24.
 *
 * Restates risc0-zkp 3.0.2 `core/hash/poseidon2/{mod,consts,rng}.rs` (un-vendored dependency of
 * /root/reference, Cargo.lock:5393-5396; reached from crates/host/src/lib.rs:137) as described in
 * SURVEY.md Appendix A.4:
 *   poseidon2_mix      : M_E; 4 x {add 24 rc; x^7 all; M_E}; 21 x {add rc to cell 0; x^7 cell 0; M_I}; 4 x full.
 *   unpadded_hash      : overwrite-mode sponge, rate 16, zero-pad the trailing partial block (or empty input).
 *   hash_pair          : 16 digest words as raw Montgomery Elems, one permutation, cells 0..7 out.
 *   Poseidon2Rng       : mix / random_elem / random_ext_elem / random_bits.
 * PARITY: permutation pinned by the 24-word KAT of SURVEY.md B.3 (tests/test_oracle_poseidon2.py);
 * sponge / RNG conventions are "parity unpinned" (recalled, medium confidence).
 */
#include "oracle.h"
#include "poseidon2_consts.h"
#include <string.h>

static fp RC[P2_ROUND_CONSTANTS_LEN];
static fp DIAG[24];
static int p2_ready = 0;

void or_poseidon2_init(void) {
  if (p2_ready) return;
  for (int i = 0; i < P2_ROUND_CONSTANTS_LEN; i++) RC[i] = fp_from_u32(P2_ROUND_CONSTANTS[i]);
  for (int i = 0; i < 24; i++) DIAG[i] = fp_from_u32(P2_M_INT_DIAG[i]);
  p2_ready = 1;
}

static inline fp sbox(fp x) {
  fp x2 = fp_mul(x, x);
  fp x4 = fp_mul(x2, x2);
  fp x6 = fp_mul(x4, x2);
  return fp_mul(x6, x);
}

/* M4 = [[5,7,1,3],[4,6,1,1],[1,3,5,7],[1,1,4,6]] on each 4-chunk, then add column sums of the chunks. */
static void m_ext(fp *s) {
  for (int c = 0; c < 24; c += 4) {
    fp a = s[c], b = s[c + 1], cc = s[c + 2], d = s[c + 3];
    fp t0 = fp_add(a, b);               /* a+b */
    fp t1 = fp_add(cc, d);              /* c+d */
    fp t2 = fp_add(fp_add(b, b), t1);   /* 2b+c+d */
    fp t3 = fp_add(fp_add(d, d), t0);   /* a+b+2d */
    fp t4 = fp_add(fp_add(t1, t1), fp_add(t1, t1)); /* 4c+4d */
    t4 = fp_add(t4, t3);                /* a+b+4c+6d */
    fp t5 = fp_add(fp_add(t0, t0), fp_add(t0, t0)); /* 4a+4b */
    t5 = fp_add(t5, t2);                /* 4a+6b+c+d */
    fp t6 = fp_add(t3, t5);             /* 5a+7b+c+3d */
    fp t7 = fp_add(t2, t4);             /* a+3b+5c+7d */
    s[c] = t6;
    s[c + 1] = t5;
    s[c + 2] = t7;
    s[c + 3] = t4;
  }
  fp sums[4] = {0, 0, 0, 0};
  for (int i = 0; i < 24; i++) sums[i & 3] = fp_add(sums[i & 3], s[i]);
  for (int i = 0; i < 24; i++) s[i] = fp_add(s[i], sums[i & 3]);
}

static void m_int(fp *s) {
  fp tot = 0;
  for (int i = 0; i < 24; i++) tot = fp_add(tot, s[i]);
  for (int i = 0; i < 24; i++) s[i] = fp_add(tot, fp_mul(DIAG[i], s[i]));
}

void or_poseidon2_mix(fp *s) {
  or_poseidon2_init();
  const fp *rc = RC;
  m_ext(s);
  for (int r = 0; r < 4; r++) {
    for (int i = 0; i < 24; i++) s[i] = sbox(fp_add(s[i], rc[i]));
    rc += 24;
    m_ext(s);
  }
  for (int r = 0; r < 21; r++) {
    s[0] = sbox(fp_add(s[0], *rc++));
    m_int(s);
  }
  for (int r = 0; r < 4; r++) {
    for (int i = 0; i < 24; i++) s[i] = sbox(fp_add(s[i], rc[i]));
    rc += 24;
    m_ext(s);
  }
}

/* ---- the same permutation on P2_LANES independent states at once (st[cell][lane]): every field operation is a loop over
 * lanes that the compiler vectorises.  Used by or_hash_rows / or_hash_fold (hal.c), which hash 16 rows / 16 node pairs per
 * call; tests/test_oracle_circuit.py checks them against the scalar sponge row by row. ---- */
#define PL P2_LANES
static inline void lv_acc(fp *restrict acc, const fp *b) {
  for (int l = 0; l < PL; l++) acc[l] = fp_add(acc[l], b[l]);
}
static inline void lv_sbox_rc(fp *io, fp rc) {
  for (int l = 0; l < PL; l++) io[l] = sbox(fp_add(io[l], rc));
}
static void m_ext_lanes(fp (*s)[PL]) {
  fp sums[4][PL];
  memset(sums, 0, sizeof sums);
  for (int c = 0; c < 24; c += 4) {
    for (int l = 0; l < PL; l++) {
      fp a = s[c][l], b = s[c + 1][l], cc = s[c + 2][l], d = s[c + 3][l];
      fp t0 = fp_add(a, b);
      fp t1 = fp_add(cc, d);
      fp t2 = fp_add(fp_add(b, b), t1);
      fp t3 = fp_add(fp_add(d, d), t0);
      fp t4 = fp_add(fp_add(t1, t1), fp_add(t1, t1));
      t4 = fp_add(t4, t3);
      fp t5 = fp_add(fp_add(t0, t0), fp_add(t0, t0));
      t5 = fp_add(t5, t2);
      fp t6 = fp_add(t3, t5);
      fp t7 = fp_add(t2, t4);
      s[c][l] = t6;
      s[c + 1][l] = t5;
      s[c + 2][l] = t7;
      s[c + 3][l] = t4;
      sums[0][l] = fp_add(sums[0][l], t6);
      sums[1][l] = fp_add(sums[1][l], t5);
      sums[2][l] = fp_add(sums[2][l], t7);
      sums[3][l] = fp_add(sums[3][l], t4);
    }
  }
  for (int i = 0; i < 24; i++)
    for (int l = 0; l < PL; l++) s[i][l] = fp_add(s[i][l], sums[i & 3][l]);
}
static void m_int_lanes(fp (*s)[PL]) {
  fp tot[PL];
  memset(tot, 0, sizeof tot);
  for (int i = 0; i < 24; i++) lv_acc(tot, s[i]);
  for (int i = 0; i < 24; i++)
    for (int l = 0; l < PL; l++) s[i][l] = fp_add(tot[l], fp_mul(DIAG[i], s[i][l]));
}
void or_poseidon2_mix_lanes(fp (*s)[P2_LANES]) {
  or_poseidon2_init();
  const fp *rc = RC;
  m_ext_lanes(s);
  for (int r = 0; r < 4; r++) {
    for (int i = 0; i < 24; i++) lv_sbox_rc(s[i], rc[i]);
    rc += 24;
    m_ext_lanes(s);
  }
  for (int r = 0; r < 21; r++) {
    lv_sbox_rc(s[0], *rc++);
    m_int_lanes(s);
  }
  for (int r = 0; r < 4; r++) {
    for (int i = 0; i < 24; i++) lv_sbox_rc(s[i], rc[i]);
    rc += 24;
    m_ext_lanes(s);
  }
}
#undef PL

/* unpadded_hash over `n` elements read with stride `stride` (stride lets hash_rows walk a column-major matrix). */
void or_hash_elems_strided(const fp *in, size_t n, size_t stride, uint32_t *digest) {
  fp st[24];
  memset(st, 0, sizeof st);
  size_t unmixed = 0;
  for (size_t i = 0; i < n; i++) {
    st[unmixed++] = in[i * stride];
    if (unmixed == 16) {
      or_poseidon2_mix(st);
      unmixed = 0;
    }
  }
  if (unmixed != 0 || n == 0) {
    for (size_t i = unmixed; i < 16; i++) st[i] = 0;
    or_poseidon2_mix(st);
  }
  memcpy(digest, st, 8 * sizeof(uint32_t));
}
void or_hash_elem_slice(const fp *in, size_t n, uint32_t *digest) { or_hash_elems_strided(in, n, 1, digest); }
/* hash_ext_elem_slice: sub-elements in order == the flat u32 view of an fp4 array. */
void or_hash_ext_elem_slice(const fp4 *in, size_t n, uint32_t *digest) { or_hash_elems_strided((const fp *)in, 4 * n, 1, digest); }
void or_hash_pair(const uint32_t *a, const uint32_t *b, uint32_t *digest) {
  fp st[24];
  memcpy(st, a, 32);
  memcpy(st + 8, b, 32);
  memset(st + 16, 0, 32);
  or_poseidon2_mix(st);
  memcpy(digest, st, 32);
}

/* ---- Poseidon2Rng (rng.rs) ---- */
void or_rng_init(or_rng *r) {
  memset(r, 0, sizeof *r);
}
void or_rng_mix(or_rng *r, const uint32_t *digest) {
  if (r->pool_used != 0) {
    or_poseidon2_mix(r->cells);
    r->pool_used = 0;
  }
  for (int i = 0; i < 8; i++) r->cells[i] = fp_add(r->cells[i], digest[i]);
  or_poseidon2_mix(r->cells);
}
fp or_rng_elem(or_rng *r) {
  if (r->pool_used == 16) {
    or_poseidon2_mix(r->cells);
    r->pool_used = 0;
  }
  return r->cells[r->pool_used++];
}
fp4 or_rng_ext(or_rng *r) {
  fp4 o;
  for (int i = 0; i < 4; i++) o.e[i] = or_rng_elem(r);
  return o;
}
uint32_t or_rng_bits(or_rng *r, unsigned bits) {
  uint32_t val = fp_to_u32(or_rng_elem(r));
  for (int i = 0; i < PP_RANDOM_BITS_EXTRA_DRAWS; i++) { /* [RECALL lo-med] */
    uint32_t nv = fp_to_u32(or_rng_elem(r));
    if (val == 0) val = nv;
  }
  return (uint32_t)(((uint64_t)1 << bits) - 1) & val;
}
