/* ORACLE (test infrastructure, not product code): transcript, Merkle trees, FRI, the segment prover driver and
 * a verifier restatement.
 *
 * Restates risc0-zkp 3.0.2 (un-vendored dependency of /root/reference, Cargo.lock:5393-5396; entered from
 * crates/host/src/lib.rs:137 and, for the verifier, crates/host/src/bin/cli.rs:103):
 *   prove/write_iop.rs, verify/read_iop.rs      -> iop_*            
 *   core/merkle.rs, prove/merkle.rs, verify/merkle.rs -> merkle_*   (A.5)
 *   prove/poly_group.rs, prove/prover.rs        -> poly_group_*, or_prove_segment (A.3, A.6)
 *   prove/fri.rs, verify/fri.rs                 -> fri_prove, fri_verify (A.10)
 *   verify/mod.rs                               -> or_verify_segment
 * and the segment-prover call sequence of risc0-circuit-rv32im 4.0.2 `prove/`.
 * PARITY: "parity unpinned" against the reference (no golden seal exists here); anchored by prove->verify
 * acceptance and single-word tamper rejection in tests/test_oracle_prover.py.
 */
#include "oracle.h"
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* circuit accessors (circuit.c) */
size_t orc_n_taps(const or_circuit *c);
size_t orc_n_regs(const or_circuit *c);
size_t orc_n_combos(const or_circuit *c);
size_t orc_tot_combo_backs(const or_circuit *c);
const uint32_t *orc_tap_group(const or_circuit *c);
const uint32_t *orc_tap_offset(const or_circuit *c);
const uint32_t *orc_tap_back(const or_circuit *c);
const uint32_t *orc_reg_group(const or_circuit *c);
const uint32_t *orc_reg_offset(const or_circuit *c);
const uint32_t *orc_reg_combo(const or_circuit *c);
const uint32_t *orc_reg_tap_pos(const or_circuit *c);
const uint32_t *orc_reg_size(const or_circuit *c);
const uint32_t *orc_combo_begin(const or_circuit *c);
const uint32_t *orc_combo_backs(const or_circuit *c);
const uint8_t *orc_info(const or_circuit *c);
const uint32_t *orc_control_root(const or_circuit *c, unsigned po2);

static double now_s(void) { return omp_get_wtime(); }

/* ---------------- WriteIOP / ReadIOP ---------------- */
typedef struct {
  uint32_t *proof;
  size_t len, cap;
  or_rng rng;
} write_iop;
static void wiop_init(write_iop *w) {
  w->cap = 1 << 16;
  w->len = 0;
  w->proof = (uint32_t *)malloc(4 * w->cap);
  or_rng_init(&w->rng);
}
static void wiop_write(write_iop *w, const uint32_t *words, size_t n) {
  if (w->len + n > w->cap) {
    while (w->len + n > w->cap) w->cap *= 2;
    w->proof = (uint32_t *)realloc(w->proof, 4 * w->cap);
  }
  memcpy(w->proof + w->len, words, 4 * n);
  w->len += n;
}
typedef struct {
  const uint32_t *proof;
  size_t len, pos;
  or_rng rng;
  int bad;
} read_iop;
static const uint32_t *riop_read(read_iop *r, size_t n) {
  if (r->pos + n > r->len) {
    r->bad = 1;
    return NULL;
  }
  const uint32_t *p = r->proof + r->pos;
  r->pos += n;
  return p;
}

/* ---------------- Merkle ---------------- */
typedef struct {
  size_t row_size, col_size, layers, top_layer, top_size;
} merkle_params;
static merkle_params merkle_params_new(size_t row_size, size_t col_size, size_t queries) {
  merkle_params p;
  p.row_size = row_size;
  p.col_size = col_size;
  p.layers = bb_log2(row_size);
  p.top_layer = pp_merkle_top_layer(p.layers, queries); /* [RECALL med] rule lives in protocol_params.h */
  p.top_size = (size_t)1 << p.top_layer;
  return p;
}
typedef struct {
  merkle_params params;
  const fp *matrix;
  uint32_t *nodes; /* 2*row_size digests */
} merkle_prover;
static void merkle_prover_new(merkle_prover *m, const fp *matrix, size_t rows, size_t cols) {
  m->params = merkle_params_new(rows, cols, OR_QUERIES);
  m->matrix = matrix;
  m->nodes = (uint32_t *)calloc(2 * rows * 8, 4);
  or_hash_rows(m->nodes + 8 * rows, matrix, rows, cols);
  for (size_t i = m->params.layers; i-- > 0;) or_hash_fold(m->nodes, (size_t)2 << i, (size_t)1 << i);
}
static const uint32_t *merkle_root(const merkle_prover *m) { return m->nodes + 8; }
static void merkle_commit(const merkle_prover *m, write_iop *iop) {
  size_t ts = m->params.top_size;
  wiop_write(iop, m->nodes + 8 * ts, 8 * ts);
  or_rng_mix(&iop->rng, merkle_root(m));
}
static void merkle_prove(const merkle_prover *m, write_iop *iop, size_t idx) {
  size_t rows = m->params.row_size, cols = m->params.col_size;
  fp *row = (fp *)malloc(4 * cols);
  or_gather_sample(row, m->matrix, idx, cols, rows);
  wiop_write(iop, row, cols);
  free(row);
  idx += rows;
  while (idx >= 2 * m->params.top_size) {
    wiop_write(iop, m->nodes + 8 * (idx ^ 1), 8);
    idx /= 2;
  }
}
static void merkle_free(merkle_prover *m) { free(m->nodes); }

typedef struct {
  merkle_params params;
  uint32_t *top; /* 2*top_size digests */
} merkle_verifier;
static int merkle_verifier_new(merkle_verifier *v, read_iop *iop, size_t rows, size_t cols) {
  v->params = merkle_params_new(rows, cols, OR_QUERIES);
  size_t ts = v->params.top_size;
  v->top = (uint32_t *)calloc(2 * ts * 8, 4);
  const uint32_t *p = riop_read(iop, 8 * ts);
  if (!p) return 1;
  for (size_t i = 0; i < 8 * ts; i++)
    if (p[i] >= BB_P) return 1; /* hash_pair asserts reduced digest words: a non-canonical word is a malleated seal */
  memcpy(v->top + 8 * ts, p, 32 * ts);
  for (size_t i = ts; i-- > 1;) or_hash_pair(v->top + 8 * (2 * i), v->top + 8 * (2 * i + 1), v->top + 8 * i);
  or_rng_mix(&iop->rng, v->top + 8);
  return 0;
}
/* returns pointer to the row (inside the seal) or NULL on failure */
static const fp *merkle_verify(const merkle_verifier *v, read_iop *iop, size_t idx) {
  if (idx >= v->params.row_size) return NULL;
  const fp *row = riop_read(iop, v->params.col_size);
  if (!row) return NULL;
  for (size_t i = 0; i < v->params.col_size; i++)
    if (row[i] >= BB_P) return NULL;
  uint32_t cur[8], tmp[8];
  or_hash_elem_slice(row, v->params.col_size, cur);
  idx += v->params.row_size;
  while (idx >= 2 * v->params.top_size) {
    size_t low = idx & 1;
    const uint32_t *other = riop_read(iop, 8);
    if (!other) return NULL;
    for (int i = 0; i < 8; i++)
      if (other[i] >= BB_P) return NULL; /* non-canonical sibling digest: reject (seal malleability) */
    idx /= 2;
    if (low) or_hash_pair(other, cur, tmp);
    else or_hash_pair(cur, other, tmp);
    memcpy(cur, tmp, 32);
  }
  if (memcmp(v->top + 8 * idx, cur, 32) != 0) return NULL;
  return row;
}

/* ---------------- PolyGroup ---------------- */
typedef struct {
  fp *coeffs;    /* [count][n], natural order after construction */
  fp *evaluated; /* [count][4n] */
  size_t count, n;
  merkle_prover merkle;
} poly_group;
/* prove/poly_group.rs PolyGroup::new -- takes ownership of coeffs (bit-reversed, zk-shifted) */
static void poly_group_new(poly_group *pg, fp *coeffs, size_t count, size_t n) {
  size_t domain = n * OR_INV_RATE;
  pg->coeffs = coeffs;
  pg->count = count;
  pg->n = n;
  pg->evaluated = (fp *)malloc(4 * count * domain);
  or_batch_expand_into_evaluate_ntt(pg->evaluated, coeffs, count, n, 2);
  or_batch_bit_reverse(coeffs, count, n);
  merkle_prover_new(&pg->merkle, pg->evaluated, domain, count);
}
static void poly_group_free(poly_group *pg) {
  free(pg->coeffs);
  free(pg->evaluated);
  merkle_free(&pg->merkle);
}
/* prove/prover.rs make_coeffs + commit_group */
static void commit_group(poly_group *pg, write_iop *iop, const fp *witness, size_t count, size_t n) {
  fp *coeffs = (fp *)malloc(4 * count * n);
  memcpy(coeffs, witness, 4 * count * n);
  or_batch_interpolate_ntt(coeffs, count, n);
  or_zk_shift(coeffs, count, n);
  poly_group_new(pg, coeffs, count, n);
  merkle_commit(&pg->merkle, iop);
}

/* Control root of a code group: the Merkle root its commitment will have in any honest seal (risc0 ships these per po2
 * as the control-ID table; verify/mod.rs check_code compares the seal's code root against it). */
void or_control_root(unsigned po2, const fp *code, size_t cols, uint32_t *root8) {
  size_t n = (size_t)1 << po2;
  poly_group pg;
  fp *coeffs = (fp *)malloc(4 * cols * n);
  memcpy(coeffs, code, 4 * cols * n);
  or_batch_interpolate_ntt(coeffs, cols, n);
  or_zk_shift(coeffs, cols, n);
  poly_group_new(&pg, coeffs, cols, n);
  memcpy(root8, merkle_root(&pg.merkle), 32);
  poly_group_free(&pg);
}

/* ---------------- small polynomial helpers (core/poly.rs) ---------------- */
static fp4 poly_eval(const fp4 *coeffs, size_t n, fp4 x) {
  fp4 tot = fp4_zero();
  for (size_t i = n; i-- > 0;) tot = fp4_add(fp4_mul(tot, x), coeffs[i]);
  return tot;
}
/* unique interpolant through (x_i, fx_i), Newton form expanded to monomial coefficients */
static void poly_interpolate(fp4 *out, const fp4 *x, const fp4 *fx, size_t size) {
  fp4 basis[18], dd[18];
  if (size > 16) abort();
  for (size_t i = 0; i < size; i++) dd[i] = fx[i];
  for (size_t j = 1; j < size; j++)
    for (size_t i = size - 1; i >= j; i--) dd[i] = fp4_mul(fp4_sub(dd[i], dd[i - 1]), fp4_inv(fp4_sub(x[i], x[i - j])));
  for (size_t i = 0; i < size; i++) out[i] = fp4_zero();
  basis[0] = fp4_one();
  size_t blen = 1;
  for (size_t j = 0; j < size; j++) {
    for (size_t i = 0; i < blen; i++) out[i] = fp4_add(out[i], fp4_mul(dd[j], basis[i]));
    /* basis *= (X - x_j) */
    basis[blen] = fp4_zero();
    for (size_t i = blen; i > 0; i--) basis[i] = fp4_sub(basis[i - 1], fp4_mul(basis[i], x[j]));
    basis[0] = fp4_neg(fp4_mul(basis[0], x[j]));
    blen++;
  }
}
/* in-place synthetic division by (X - z); returns the remainder */
static fp4 poly_divide(fp4 *p, size_t n, fp4 z) {
  fp4 cur = fp4_zero();
  for (size_t i = n; i-- > 0;) {
    fp4 next = fp4_add(fp4_mul(z, cur), p[i]);
    p[i] = cur;
    cur = next;
  }
  return cur;
}

/* ---------------- FRI (prove/fri.rs) ---------------- */
typedef struct {
  size_t domain;
  fp *coeffs, *evaluated;
  merkle_prover merkle;
} fri_round;

typedef void (*inner_prove_fn)(void *ctx, write_iop *iop, size_t idx);

static void fri_prove(write_iop *iop, const fp *coeffs_in, size_t size0, inner_prove_fn inner, void *ctx) {
  size_t orig_domain = size0 * OR_INV_RATE;
  fri_round rounds[8];
  size_t nr = 0;
  size_t size = size0;
  fp *coeffs = (fp *)malloc(16 * size);
  memcpy(coeffs, coeffs_in, 16 * size);
  while (size > OR_FRI_MIN_DEGREE) {
    fri_round *r = &rounds[nr++];
    r->domain = size * OR_INV_RATE;
    r->evaluated = (fp *)malloc(16 * r->domain);
    or_batch_expand_into_evaluate_ntt(r->evaluated, coeffs, 4, size, 2);
    merkle_prover_new(&r->merkle, r->evaluated, r->domain / OR_FRI_FOLD, OR_FRI_FOLD * 4);
    merkle_commit(&r->merkle, iop);
    fp4 fold_mix = or_rng_ext(&iop->rng);
    r->coeffs = (fp *)malloc(16 * size / OR_FRI_FOLD);
    or_fri_fold(r->coeffs, coeffs, &fold_mix, size / OR_FRI_FOLD);
    free(coeffs);
    size /= OR_FRI_FOLD;
    coeffs = (fp *)malloc(16 * size);
    memcpy(coeffs, r->coeffs, 16 * size);
  }
  /* final coefficients in natural order */
  or_batch_bit_reverse(coeffs, 4, size);
  wiop_write(iop, coeffs, 4 * size);
  uint32_t dg[8];
  or_hash_elem_slice(coeffs, 4 * size, dg);
  or_rng_mix(&iop->rng, dg);
  free(coeffs);
  for (size_t q = 0; q < OR_QUERIES; q++) {
    size_t pos = or_rng_bits(&iop->rng, bb_log2(orig_domain));
    inner(ctx, iop, pos);
    for (size_t k = 0; k < nr; k++) {
      size_t group = pos % (rounds[k].domain / OR_FRI_FOLD);
      merkle_prove(&rounds[k].merkle, iop, group);
      pos = group;
    }
  }
  for (size_t k = 0; k < nr; k++) {
    free(rounds[k].coeffs);
    free(rounds[k].evaluated);
    merkle_free(&rounds[k].merkle);
  }
}

/* ---------------- segment prover ---------------- */
typedef struct {
  poly_group *groups[3];
  poly_group *check;
} inner_ctx;
static void inner_prove(void *vctx, write_iop *iop, size_t idx) {
  inner_ctx *c = (inner_ctx *)vctx;
  for (int k = 0; k < 3; k++) merkle_prove(&c->groups[PP_QUERY_ORDER[k]]->merkle, iop, idx); /* [RECALL med] */
  merkle_prove(&c->check->merkle, iop, idx);
}

static const char PROOF_SYSTEM_INFO[PP_INFO_BYTES + 1] = PP_PROOF_SYSTEM_INFO; /* [RECALL med] protocol_params.h */
static void hash_info(const uint8_t *info, uint32_t *dg) {
  fp e[16];
  for (int i = 0; i < 16; i++) e[i] = fp_from_u32(info[i]);
  or_hash_elem_slice(e, 16, dg);
}

enum { G_ACCUM = 0, G_CODE = 1, G_DATA = 2 };

size_t or_prove_segment(const or_circuit *c, unsigned po2, const fp *code, const fp *data, const fp *globals,
                        const fp *accum_noise, size_t zk_rows, uint32_t **seal, double *tm) {
  double t0 = now_s(), t;
  size_t n = (size_t)1 << po2, domain = 4 * n;
  size_t wa = or_circuit_group_size(c, G_ACCUM), wc = or_circuit_group_size(c, G_CODE), wd = or_circuit_group_size(c, G_DATA);
  size_t gsz = or_circuit_globals_size(c), msz = or_circuit_mix_size(c);
  write_iop iop;
  wiop_init(&iop);
  uint32_t dg[8];
  hash_info((const uint8_t *)PROOF_SYSTEM_INFO, dg);
  or_rng_mix(&iop.rng, dg);
  hash_info(orc_info(c), dg);
  or_rng_mix(&iop.rng, dg);
  /* header = globals || po2 */
  fp *header = (fp *)malloc(4 * (gsz + 1));
  memcpy(header, globals, 4 * gsz);
  header[gsz] = fp_from_u32(po2);
  or_hash_elem_slice(header, gsz + 1, dg);
  or_rng_mix(&iop.rng, dg);
  wiop_write(&iop, header, gsz + 1);
  free(header);

  poly_group pg[3], check_pg;
  t = now_s();
  commit_group(&pg[G_CODE], &iop, code, wc, n);
  commit_group(&pg[G_DATA], &iop, data, wd, n);
  if (tm) tm[0] = now_s() - t;
  fp *mix = (fp *)malloc(4 * (msz + 1));
  for (size_t i = 0; i < msz; i++) mix[i] = or_rng_elem(&iop.rng);
  t = now_s();
  fp *accum = (fp *)malloc(4 * wa * n);
  or_accumulate(c, po2, accum, code, data, mix, accum_noise, zk_rows);
  if (tm) tm[1] = now_s() - t;
  t = now_s();
  commit_group(&pg[G_ACCUM], &iop, accum, wa, n);
  free(accum);
  if (tm) tm[0] += now_s() - t;

  /* ---- finalize (prove/prover.rs) ---- */
  fp4 poly_mix = or_rng_ext(&iop.rng);
  fp *check = (fp *)malloc(16 * domain);
  const fp *evals[3] = {pg[0].evaluated, pg[1].evaluated, pg[2].evaluated};
  t = now_s();
  or_eval_check(c, check, evals, globals, mix, &poly_mix, po2);
  if (tm) tm[2] = now_s() - t;
  t = now_s();
  or_batch_interpolate_ntt(check, 4, domain);
  /* the bit-reversed length-4N coefficient array, cut in 4 quarters per component, is the 16 x N check group */
  poly_group_new(&check_pg, check, OR_CHECK_SIZE, n);
  merkle_commit(&check_pg.merkle, &iop);
  if (tm) tm[3] = now_s() - t;

  t = now_s();
  fp4 z = or_rng_ext(&iop.rng);
  fp back_one = bb_rou_rev(po2);
  size_t n_taps = orc_n_taps(c), n_regs = orc_n_regs(c), n_combos = orc_n_combos(c);
  const uint32_t *tg = orc_tap_group(c), *to = orc_tap_offset(c), *tb = orc_tap_back(c);
  fp4 *all_xs = (fp4 *)malloc(sizeof(fp4) * (n_taps + 1));
  fp4 *eval_u = (fp4 *)malloc(sizeof(fp4) * (n_taps + 1));
  uint32_t *which = (uint32_t *)malloc(4 * (n_taps + 16));
  {
    size_t pos = 0;
    for (int g = 0; g < 3; g++) {
      size_t start = pos;
      while (pos < n_taps && tg[pos] == (uint32_t)g) {
        which[pos - start] = to[pos];
        all_xs[pos] = fp4_mul_fp(z, fp_pow(back_one, tb[pos]));
        pos++;
      }
      or_batch_evaluate_any(pg[g].coeffs, pg[g].count, n, which, all_xs + start, pos - start, eval_u + start);
    }
  }
  fp4 *coeff_u = (fp4 *)malloc(sizeof(fp4) * (n_taps + OR_CHECK_SIZE));
  {
    const uint32_t *rp = orc_reg_tap_pos(c), *rs = orc_reg_size(c);
    for (size_t r = 0; r < n_regs; r++) poly_interpolate(coeff_u + rp[r], all_xs + rp[r], eval_u + rp[r], rs[r]);
  }
  fp4 z_pow = fp4_pow(z, OR_EXT_SIZE);
  {
    fp4 xs[OR_CHECK_SIZE];
    for (uint32_t i = 0; i < OR_CHECK_SIZE; i++) {
      which[i] = i;
      xs[i] = z_pow;
    }
    or_batch_evaluate_any(check_pg.coeffs, OR_CHECK_SIZE, n, which, xs, OR_CHECK_SIZE, coeff_u + n_taps);
  }
  wiop_write(&iop, (const uint32_t *)coeff_u, 4 * (n_taps + OR_CHECK_SIZE));
  or_hash_ext_elem_slice(coeff_u, n_taps + OR_CHECK_SIZE, dg);
  or_rng_mix(&iop.rng, dg);
  if (tm) tm[4] = now_s() - t;

  t = now_s();
  fp4 mixv = or_rng_ext(&iop.rng);
  fp4 *combos = (fp4 *)calloc(n * (n_combos + 1), sizeof(fp4));
  fp4 cur_mix = fp4_one();
  {
    const uint32_t *rg = orc_reg_group(c), *rc = orc_reg_combo(c);
    size_t r = 0;
    for (int g = 0; g < 3; g++) {
      size_t gs = or_circuit_group_size(c, g);
      uint32_t *cw = (uint32_t *)malloc(4 * (gs + 1));
      size_t k = 0;
      while (r < n_regs && rg[r] == (uint32_t)g) cw[k++] = rc[r++];
      if (k != gs) { fprintf(stderr, "oracle: group %d has %zu regs for %zu columns\n", g, k, gs); abort(); }
      or_mix_poly_coeffs(combos, &cur_mix, &mixv, pg[g].coeffs, cw, gs, n);
      cur_mix = fp4_mul(cur_mix, fp4_pow(mixv, gs));
      free(cw);
    }
    uint32_t cw[OR_CHECK_SIZE];
    for (int i = 0; i < OR_CHECK_SIZE; i++) cw[i] = (uint32_t)n_combos;
    or_mix_poly_coeffs(combos, &cur_mix, &mixv, check_pg.coeffs, cw, OR_CHECK_SIZE, n);
  }
  {
    const uint32_t *rc = orc_reg_combo(c), *rs = orc_reg_size(c);
    size_t cur_pos = 0;
    fp4 cur = fp4_one();
    for (size_t r = 0; r < n_regs; r++) {
      for (size_t i = 0; i < rs[r]; i++) {
        fp4 *p = &combos[n * rc[r] + i];
        *p = fp4_sub(*p, fp4_mul(cur, coeff_u[cur_pos + i]));
      }
      cur = fp4_mul(cur, mixv);
      cur_pos += rs[r];
    }
    for (int i = 0; i < OR_CHECK_SIZE; i++) {
      fp4 *p = &combos[n * n_combos];
      *p = fp4_sub(*p, fp4_mul(cur, coeff_u[cur_pos]));
      cur_pos++;
      cur = fp4_mul(cur, mixv);
    }
    const uint32_t *cb = orc_combo_begin(c), *cbk = orc_combo_backs(c);
    int bad = 0;
#pragma omp parallel for schedule(dynamic) reduction(| : bad)
    for (size_t k = 0; k <= n_combos; k++) {
      if (k < n_combos) {
        for (size_t j = cb[k]; j < cb[k + 1]; j++) {
          fp4 rem = poly_divide(combos + k * n, n, fp4_mul_fp(z, fp_pow(back_one, cbk[j])));
          if (!fp4_eq(rem, fp4_zero())) bad |= 1;
        }
      } else {
        fp4 rem = poly_divide(combos + n_combos * n, n, z_pow);
        if (!fp4_eq(rem, fp4_zero())) bad |= 1;
      }
    }
    if (bad) fprintf(stderr, "oracle: nonzero remainder in DEEP division\n");
  }
  fp *final_coeffs = (fp *)malloc(16 * n);
  or_eltwise_sum_extelem(final_coeffs, combos, n, n_combos + 1);
  free(combos);
  or_batch_bit_reverse(final_coeffs, 4, n);
  if (tm) tm[5] = now_s() - t;

  t = now_s();
  inner_ctx ictx = {{&pg[0], &pg[1], &pg[2]}, &check_pg};
  fri_prove(&iop, final_coeffs, n, inner_prove, &ictx);
  if (tm) tm[6] = now_s() - t;
  free(final_coeffs);
  free(all_xs); free(eval_u); free(which); free(coeff_u); free(mix);
  for (int g = 0; g < 3; g++) poly_group_free(&pg[g]);
  poly_group_free(&check_pg);
  *seal = iop.proof;
  if (tm) tm[7] = now_s() - t0;
  return iop.len;
}

/* ---------------- verifier (verify/mod.rs, verify/fri.rs) ---------------- */
typedef struct {
  size_t domain;
  merkle_verifier merkle;
  fp4 mix;
} vround;

static void ext_interpolate_ntt16(fp4 *io) { /* core/ntt.rs interpolate_ntt on 16 ext values */
  size_t n = 16;
  for (unsigned s = 4; s >= 1; s--) {
    size_t half = (size_t)1 << (s - 1);
    fp step = bb_rou_rev(s);
    for (size_t blk = 0; blk < n; blk += 2 * half) {
      fp cur = FP_ONE;
      for (size_t i = 0; i < half; i++) {
        fp4 a = io[blk + i], b = io[blk + i + half];
        io[blk + i] = fp4_add(a, b);
        io[blk + i + half] = fp4_mul_fp(fp4_sub(a, b), cur);
        cur = fp_mul(cur, step);
      }
    }
  }
  fp norm = fp_inv(fp_from_u32(16));
  for (size_t i = 0; i < n; i++) io[i] = fp4_mul_fp(io[i], norm);
}

int or_verify_segment(const or_circuit *c, const uint32_t *seal, size_t seal_words) {
  read_iop iop;
  iop.proof = seal;
  iop.len = seal_words;
  iop.pos = 0;
  iop.bad = 0;
  or_rng_init(&iop.rng);
  size_t wa = or_circuit_group_size(c, G_ACCUM), wc = or_circuit_group_size(c, G_CODE), wd = or_circuit_group_size(c, G_DATA);
  size_t gsz = or_circuit_globals_size(c), msz = or_circuit_mix_size(c);
  uint32_t dg[8];
  hash_info((const uint8_t *)PROOF_SYSTEM_INFO, dg);
  or_rng_mix(&iop.rng, dg);
  hash_info(orc_info(c), dg);
  or_rng_mix(&iop.rng, dg);
  const fp *header = riop_read(&iop, gsz + 1);
  if (!header) return 1;
  for (size_t i = 0; i <= gsz; i++)
    if (header[i] >= BB_P) return 2;
  or_hash_elem_slice(header, gsz + 1, dg);
  or_rng_mix(&iop.rng, dg);
  const fp *globals = header;
  uint32_t po2 = fp_to_u32(header[gsz]);
  if (po2 > 24 || po2 < 1) return 3;
  size_t n = (size_t)1 << po2, domain = 4 * n;
  int rc_ = 0;
  merkle_verifier mv[3], mcheck;
  memset(mv, 0, sizeof mv);
  memset(&mcheck, 0, sizeof mcheck);
  vround rounds[8];
  size_t nrounds = 0;
  fp *mix = NULL;
  fp4 *eval_u = NULL, *combo_u = NULL, *tap_mix_pows = NULL, *tot = NULL, *poly_buf = NULL;
#define FAIL(code) do { rc_ = (code); goto done; } while (0)
  if (merkle_verifier_new(&mv[G_CODE], &iop, domain, wc)) FAIL(4);
  { /* verify/mod.rs check_code(po2, code_merkle.root()): the seal must commit to THIS program's code columns */
    const uint32_t *want = orc_control_root(c, po2);
    if (!want) FAIL(15);
    if (memcmp(want, mv[G_CODE].top + 8, 32) != 0) FAIL(14);
  }
  if (merkle_verifier_new(&mv[G_DATA], &iop, domain, wd)) FAIL(4);
  mix = (fp *)malloc(4 * (msz + 1));
  for (size_t i = 0; i < msz; i++) mix[i] = or_rng_elem(&iop.rng);
  if (merkle_verifier_new(&mv[G_ACCUM], &iop, domain, wa)) FAIL(4);
  fp4 poly_mix = or_rng_ext(&iop.rng);
  if (merkle_verifier_new(&mcheck, &iop, domain, OR_CHECK_SIZE)) FAIL(4);
  fp4 z = or_rng_ext(&iop.rng);
  fp back_one = bb_rou_rev(po2);
  size_t n_taps = orc_n_taps(c), n_regs = orc_n_regs(c), n_combos = orc_n_combos(c), tcb = orc_tot_combo_backs(c);
  const fp4 *coeff_u = (const fp4 *)riop_read(&iop, 4 * (n_taps + OR_CHECK_SIZE));
  if (!coeff_u) FAIL(5);
  for (size_t i = 0; i < 4 * (n_taps + OR_CHECK_SIZE); i++)
    if (((const fp *)coeff_u)[i] >= BB_P) FAIL(2);
  or_hash_ext_elem_slice(coeff_u, n_taps + OR_CHECK_SIZE, dg);
  or_rng_mix(&iop.rng, dg);
  const uint32_t *tb = orc_tap_back(c), *rp = orc_reg_tap_pos(c), *rs = orc_reg_size(c), *rcb = orc_reg_combo(c),
                 *rg = orc_reg_group(c), *ro = orc_reg_offset(c), *cb = orc_combo_begin(c), *cbk = orc_combo_backs(c);
  eval_u = (fp4 *)malloc(sizeof(fp4) * (n_taps + 1));
  for (size_t r = 0; r < n_regs; r++)
    for (size_t i = 0; i < rs[r]; i++) {
      fp4 x = fp4_mul_fp(z, fp_pow(back_one, tb[rp[r] + i]));
      eval_u[rp[r] + i] = poly_eval(coeff_u + rp[r], rs[r], x);
    }
  fp4 result = or_poly_ext(c, &poly_mix, eval_u, globals, mix);
  {
    const int *remap = PP_CHECK_REMAP; /* [RECALL med] */
    fp4 check = fp4_zero();
    for (int i = 0; i < 4; i++) {
      fp4 zi = fp4_pow(z, (uint64_t)i);
      for (int k = 0; k < 4; k++) {
        fp4 basis = fp4_zero();
        basis.e[k] = FP_ONE;
        check = fp4_add(check, fp4_mul(fp4_mul(coeff_u[n_taps + remap[i] + 4 * k], zi), basis));
      }
    }
    fp4 three_z = fp4_mul_fp(z, fp_from_u32(3));
    check = fp4_mul(check, fp4_sub(fp4_pow(three_z, n), fp4_one()));
    if (!fp4_eq(check, result)) FAIL(6);
  }
  fp4 mixv = or_rng_ext(&iop.rng);
  combo_u = (fp4 *)calloc(tcb + 1, sizeof(fp4));
  tap_mix_pows = (fp4 *)malloc(sizeof(fp4) * (n_regs + OR_CHECK_SIZE));
  {
    fp4 cur = fp4_one();
    size_t cur_pos = 0;
    for (size_t r = 0; r < n_regs; r++) {
      for (size_t i = 0; i < rs[r]; i++)
        combo_u[cb[rcb[r]] + i] = fp4_add(combo_u[cb[rcb[r]] + i], fp4_mul(cur, coeff_u[cur_pos + i]));
      tap_mix_pows[r] = cur;
      cur = fp4_mul(cur, mixv);
      cur_pos += rs[r];
    }
    for (int i = 0; i < OR_CHECK_SIZE; i++) {
      combo_u[tcb] = fp4_add(combo_u[tcb], fp4_mul(cur, coeff_u[cur_pos]));
      cur_pos++;
      tap_mix_pows[n_regs + i] = cur;
      cur = fp4_mul(cur, mixv);
    }
  }
  /* fri_verify */
  {
    size_t degree = n, orig_domain = domain, dom = domain;
    while (degree > OR_FRI_MIN_DEGREE) {
      vround *r = &rounds[nrounds];
      r->domain = dom / OR_FRI_FOLD;
      memset(&r->merkle, 0, sizeof r->merkle);
      nrounds++;
      if (merkle_verifier_new(&r->merkle, &iop, r->domain, OR_FRI_FOLD * 4)) FAIL(7);
      r->mix = or_rng_ext(&iop.rng);
      dom /= OR_FRI_FOLD;
      degree /= OR_FRI_FOLD;
    }
    const fp *final_coeffs = riop_read(&iop, 4 * degree);
    if (!final_coeffs) FAIL(8);
    for (size_t i = 0; i < 4 * degree; i++)
      if (final_coeffs[i] >= BB_P) FAIL(2);
    or_hash_elem_slice(final_coeffs, 4 * degree, dg);
    or_rng_mix(&iop.rng, dg);
    fp gen = bb_rou_fwd(bb_log2(dom));
    fp gen0 = bb_rou_fwd(bb_log2(orig_domain));
    poly_buf = (fp4 *)malloc(sizeof(fp4) * degree);
    for (size_t i = 0; i < degree; i++)
      for (int j = 0; j < 4; j++) poly_buf[i].e[j] = final_coeffs[j * degree + i];
    tot = (fp4 *)malloc(sizeof(fp4) * (n_combos + 1));
    for (size_t q = 0; q < OR_QUERIES; q++) {
      size_t pos = or_rng_bits(&iop.rng, bb_log2(orig_domain));
      /* inner: open all groups at pos and form the DEEP quotient value (verify/mod.rs fri_eval_taps) */
      fp4 x = fp4_from_fp(fp_pow(gen0, pos));
      const fp *rows[3];
      for (int k = 0; k < 3; k++) { /* [RECALL med] opening order: protocol_params.h PP_QUERY_ORDER */
        int g = PP_QUERY_ORDER[k];
        rows[g] = merkle_verify(&mv[g], &iop, pos);
        if (!rows[g]) FAIL(9);
      }
      const fp *check_row = merkle_verify(&mcheck, &iop, pos);
      if (!check_row) FAIL(9);
      for (size_t k = 0; k <= n_combos; k++) tot[k] = fp4_zero();
      for (size_t r = 0; r < n_regs; r++) tot[rcb[r]] = fp4_add(tot[rcb[r]], fp4_mul_fp(tap_mix_pows[r], rows[rg[r]][ro[r]]));
      for (int i = 0; i < OR_CHECK_SIZE; i++)
        tot[n_combos] = fp4_add(tot[n_combos], fp4_mul_fp(tap_mix_pows[n_regs + i], check_row[i]));
      fp4 goal = fp4_zero();
      for (size_t k = 0; k < n_combos; k++) {
        fp4 num = fp4_sub(tot[k], poly_eval(combo_u + cb[k], cb[k + 1] - cb[k], x));
        fp4 div = fp4_one();
        for (size_t j = cb[k]; j < cb[k + 1]; j++) div = fp4_mul(div, fp4_sub(x, fp4_mul_fp(z, fp_pow(back_one, cbk[j]))));
        goal = fp4_add(goal, fp4_mul(num, fp4_inv(div)));
      }
      {
        fp4 num = fp4_sub(tot[n_combos], combo_u[tcb]);
        fp4 div = fp4_sub(x, fp4_pow(z, OR_INV_RATE));
        goal = fp4_add(goal, fp4_mul(num, fp4_inv(div)));
      }
      for (size_t k = 0; k < nrounds; k++) {
        vround *r = &rounds[k];
        size_t quot = pos / r->domain, group = pos % r->domain;
        const fp *drow = merkle_verify(&r->merkle, &iop, group);
        if (!drow) FAIL(10);
        fp4 data_ext[16];
        for (int i = 0; i < 16; i++)
          for (int j = 0; j < 4; j++) data_ext[i].e[j] = drow[j * 16 + i];
        if (!fp4_eq(data_ext[quot], goal)) FAIL(11);
        fp inv_wk = fp_pow(bb_rou_rev(bb_log2(OR_FRI_FOLD * r->domain)), group);
        ext_interpolate_ntt16(data_ext);
        /* bit-reverse to natural order */
        for (uint32_t i = 0; i < 16; i++) {
          uint32_t j = bb_bit_rev(i, 4);
          if (i < j) {
            fp4 tmp = data_ext[i];
            data_ext[i] = data_ext[j];
            data_ext[j] = tmp;
          }
        }
        goal = poly_eval(data_ext, 16, fp4_mul_fp(r->mix, inv_wk));
        pos = group;
      }
      fp4 fx = poly_eval(poly_buf, degree, fp4_from_fp(fp_pow(gen, pos)));
      if (!fp4_eq(fx, goal)) FAIL(12);
    }
  }
  if (iop.pos != iop.len) FAIL(13); /* verify_complete */
done:
  for (int g = 0; g < 3; g++) free(mv[g].top);
  free(mcheck.top);
  for (size_t k = 0; k < nrounds; k++) free(rounds[k].merkle.top);
  free(mix); free(eval_u); free(combo_u); free(tap_mix_pows); free(tot); free(poly_buf);
  return rc_;
}
