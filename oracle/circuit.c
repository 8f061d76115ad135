/* ORACLE (test infrastructure, not product code): circuit description, constraint-program interpreter,
 * synthetic witness generation and `accumulate`.
 *
 * Restates the *shape* of risc0-zkp 3.0.2 `taps.rs` (TapSet: taps sorted by (group, offset, back), registers,
 * combos) and `adapter.rs` (PolyExtStep program: Const, ConstExt, Get, GetGlobal, Add, Sub, Mul, True, AndEqz,
 * AndCond; MixState{tot, mul}) per SURVEY.md Appendix A.7, and of the generated CircuitHal::eval_check of
 * risc0-circuit-rv32im 4.0.2 (un-vendored, Cargo.lock:5320-5323): evaluate the program on every point of the
 * 4N coset x = 3*w^i with Get(tap) = group[col][(i - 4*back) mod 4N] and write tot / (x^N - 1).
 * The real rv32im program/taps are a generated artifact that is not available here: the engine is
 * generic over a "circuit blob" (format in DESIGN.md) and is exercised with synthetic circuits whose witness
 * generator (`or_witgen`, `or_accumulate`) lives here too.  PARITY: unpinned against the reference.
 */
#include "oracle.h"
#include <omp.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define BLOB_MAGIC 0x435a3242u
enum { OP_CONST = 0, OP_CONST_EXT, OP_GET, OP_GET_GLOBAL, OP_ADD, OP_SUB, OP_MUL, OP_TRUE, OP_AND_EQZ, OP_AND_COND };

typedef struct {
  uint32_t op, a, b, c, d;
} step_t;

struct or_circuit {
  size_t group_size[3], globals_size, mix_size;
  size_t n_taps, n_steps, ret;
  uint8_t info[16];
  /* taps */
  uint32_t *tap_group, *tap_offset, *tap_back;
  /* registers */
  size_t n_regs;
  uint32_t *reg_group, *reg_offset, *reg_combo, *reg_tap_pos, *reg_size;
  /* combos */
  size_t n_combos, tot_combo_backs;
  uint32_t *combo_begin, *combo_backs;
  /* program */
  step_t *steps;
  size_t n_fp, n_mix;
  uint8_t *fp_is_ext;    /* per fp var */
  uint32_t *step_out;    /* var index produced by each step */
  uint32_t *mix_exp;     /* exponent of poly_mix carried by each mix var */
  size_t max_exp;
  /* synthetic witness description */
  size_t n_wit_steps, n_wit_stores, n_accum, n_pins;
  step_t *wit_steps;
  uint32_t *wit_store_col, *wit_store_var;
  uint32_t *accum_col;   /* data column feeding accumulator k */
  uint32_t *pin_col, *pin_global;
  uint32_t *col_kind;    /* per data column: 0 free, 1 free bit, 2 derived, 3 | idx << 8 | size << 16 | group << 24 one-hot member */
  size_t n_wit_fp;
  /* control roots (risc0: control IDs): expected Merkle root of the code group per po2 (verify/mod.rs check_code) */
  uint32_t control_root[32][8];
  uint8_t has_control_root[32];
};
void or_circuit_set_control_root(or_circuit *c, unsigned po2, const uint32_t *root8) {
  if (po2 >= 32) return;
  memcpy(c->control_root[po2], root8, 32);
  c->has_control_root[po2] = 1;
}
const uint32_t *orc_control_root(const or_circuit *c, unsigned po2) {
  return (po2 < 32 && c->has_control_root[po2]) ? c->control_root[po2] : NULL;
}

size_t or_circuit_group_size(const or_circuit *c, int g) { return c->group_size[g]; }
size_t or_circuit_globals_size(const or_circuit *c) { return c->globals_size; }
size_t or_circuit_mix_size(const or_circuit *c) { return c->mix_size; }

/* accessors used by zkp.c */
size_t orc_n_taps(const or_circuit *c) { return c->n_taps; }
size_t orc_n_regs(const or_circuit *c) { return c->n_regs; }
size_t orc_n_combos(const or_circuit *c) { return c->n_combos; }
size_t orc_tot_combo_backs(const or_circuit *c) { return c->tot_combo_backs; }
const uint32_t *orc_tap_group(const or_circuit *c) { return c->tap_group; }
const uint32_t *orc_tap_offset(const or_circuit *c) { return c->tap_offset; }
const uint32_t *orc_tap_back(const or_circuit *c) { return c->tap_back; }
const uint32_t *orc_reg_group(const or_circuit *c) { return c->reg_group; }
const uint32_t *orc_reg_offset(const or_circuit *c) { return c->reg_offset; }
const uint32_t *orc_reg_combo(const or_circuit *c) { return c->reg_combo; }
const uint32_t *orc_reg_tap_pos(const or_circuit *c) { return c->reg_tap_pos; }
const uint32_t *orc_reg_size(const or_circuit *c) { return c->reg_size; }
const uint32_t *orc_combo_begin(const or_circuit *c) { return c->combo_begin; }
const uint32_t *orc_combo_backs(const or_circuit *c) { return c->combo_backs; }
const uint8_t *orc_info(const or_circuit *c) { return c->info; }

static int cmp_backs(const uint32_t *a, size_t na, const uint32_t *b, size_t nb) {
  size_t n = na < nb ? na : nb;
  for (size_t i = 0; i < n; i++)
    if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
  return na == nb ? 0 : (na < nb ? -1 : 1);
}

or_circuit *or_circuit_load(const uint32_t *blob, size_t words) {
  if (words < 24 || blob[0] != BLOB_MAGIC || blob[1] != 1) return NULL;
  or_circuit *c = (or_circuit *)calloc(1, sizeof *c);
  for (int g = 0; g < 3; g++) c->group_size[g] = blob[2 + g];
  c->globals_size = blob[5];
  c->mix_size = blob[6];
  c->n_taps = blob[7];
  c->n_steps = blob[8];
  c->ret = blob[9];
  memcpy(c->info, blob + 10, 16);
  c->n_wit_steps = blob[14];
  c->n_wit_stores = blob[15];
  c->n_accum = blob[16];
  c->n_pins = blob[17];
  size_t pos = 24;
  size_t need = pos + c->n_taps + 5 * c->n_steps + 5 * c->n_wit_steps + 2 * c->n_wit_stores + c->n_accum + 2 * c->n_pins +
                c->group_size[2];
  if (words < need) {
    free(c);
    return NULL;
  }
  c->tap_group = (uint32_t *)malloc(4 * c->n_taps);
  c->tap_offset = (uint32_t *)malloc(4 * c->n_taps);
  c->tap_back = (uint32_t *)malloc(4 * c->n_taps);
  for (size_t i = 0; i < c->n_taps; i++) {
    uint32_t w = blob[pos++];
    c->tap_offset[i] = w & 0xffff;
    c->tap_back[i] = (w >> 16) & 0xff;
    c->tap_group[i] = w >> 24;
  }
  c->steps = (step_t *)malloc(sizeof(step_t) * (c->n_steps ? c->n_steps : 1));
  memcpy(c->steps, blob + pos, 20 * c->n_steps);
  pos += 5 * c->n_steps;
  c->wit_steps = (step_t *)malloc(sizeof(step_t) * (c->n_wit_steps ? c->n_wit_steps : 1));
  memcpy(c->wit_steps, blob + pos, 20 * c->n_wit_steps);
  pos += 5 * c->n_wit_steps;
  c->wit_store_col = (uint32_t *)malloc(4 * (c->n_wit_stores + 1));
  c->wit_store_var = (uint32_t *)malloc(4 * (c->n_wit_stores + 1));
  for (size_t i = 0; i < c->n_wit_stores; i++) {
    c->wit_store_col[i] = blob[pos++];
    c->wit_store_var[i] = blob[pos++];
  }
  c->accum_col = (uint32_t *)malloc(4 * (c->n_accum + 1));
  for (size_t i = 0; i < c->n_accum; i++) c->accum_col[i] = blob[pos++];
  c->pin_col = (uint32_t *)malloc(4 * (c->n_pins + 1));
  c->pin_global = (uint32_t *)malloc(4 * (c->n_pins + 1));
  for (size_t i = 0; i < c->n_pins; i++) {
    c->pin_col[i] = blob[pos++];
    c->pin_global[i] = blob[pos++];
  }
  c->col_kind = (uint32_t *)malloc(4 * (c->group_size[2] + 1));
  for (size_t i = 0; i < c->group_size[2]; i++) c->col_kind[i] = blob[pos++];

  /* registers: runs of taps sharing (group, offset) -- taps.rs TapSet::regs() */
  c->reg_group = (uint32_t *)malloc(4 * (c->n_taps + 1));
  c->reg_offset = (uint32_t *)malloc(4 * (c->n_taps + 1));
  c->reg_combo = (uint32_t *)malloc(4 * (c->n_taps + 1));
  c->reg_tap_pos = (uint32_t *)malloc(4 * (c->n_taps + 1));
  c->reg_size = (uint32_t *)malloc(4 * (c->n_taps + 1));
  size_t nr = 0;
  for (size_t i = 0; i < c->n_taps;) {
    size_t j = i;
    while (j < c->n_taps && c->tap_group[j] == c->tap_group[i] && c->tap_offset[j] == c->tap_offset[i]) j++;
    c->reg_group[nr] = c->tap_group[i];
    c->reg_offset[nr] = c->tap_offset[i];
    c->reg_tap_pos[nr] = (uint32_t)i;
    c->reg_size[nr] = (uint32_t)(j - i);
    nr++;
    i = j;
  }
  c->n_regs = nr;
  /* combos: the distinct back-lists, numbered in lexicographic order (combo numbering) */
  uint32_t *uniq = (uint32_t *)malloc(4 * (nr + 1)); /* representative reg per combo */
  size_t nc = 0;
  for (size_t r = 0; r < nr; r++) {
    const uint32_t *b = c->tap_back + c->reg_tap_pos[r];
    size_t k;
    for (k = 0; k < nc; k++)
      if (cmp_backs(b, c->reg_size[r], c->tap_back + c->reg_tap_pos[uniq[k]], c->reg_size[uniq[k]]) == 0) break;
    if (k == nc) uniq[nc++] = (uint32_t)r;
  }
  for (size_t i = 1; i < nc; i++) { /* insertion sort, lexicographic */
    uint32_t v = uniq[i];
    size_t j = i;
    while (j > 0 && cmp_backs(c->tap_back + c->reg_tap_pos[v], c->reg_size[v], c->tap_back + c->reg_tap_pos[uniq[j - 1]],
                              c->reg_size[uniq[j - 1]]) < 0) {
      uniq[j] = uniq[j - 1];
      j--;
    }
    uniq[j] = v;
  }
  c->n_combos = nc;
  c->combo_begin = (uint32_t *)malloc(4 * (nc + 1));
  size_t tot = 0;
  for (size_t k = 0; k < nc; k++) {
    c->combo_begin[k] = (uint32_t)tot;
    tot += c->reg_size[uniq[k]];
  }
  c->combo_begin[nc] = (uint32_t)tot;
  c->tot_combo_backs = tot;
  c->combo_backs = (uint32_t *)malloc(4 * (tot + 1));
  for (size_t k = 0; k < nc; k++)
    memcpy(c->combo_backs + c->combo_begin[k], c->tap_back + c->reg_tap_pos[uniq[k]], 4 * c->reg_size[uniq[k]]);
  for (size_t r = 0; r < nr; r++) {
    const uint32_t *b = c->tap_back + c->reg_tap_pos[r];
    for (size_t k = 0; k < nc; k++)
      if (cmp_backs(b, c->reg_size[r], c->combo_backs + c->combo_begin[k], c->combo_begin[k + 1] - c->combo_begin[k]) == 0)
        c->reg_combo[r] = (uint32_t)k;
  }
  free(uniq);

  /* static typing of the program */
  c->fp_is_ext = (uint8_t *)calloc(c->n_steps + 1, 1);
  c->step_out = (uint32_t *)calloc(c->n_steps + 1, 4);
  c->mix_exp = (uint32_t *)calloc(c->n_steps + 1, 4);
  size_t nf = 0, nm = 0;
  for (size_t i = 0; i < c->n_steps; i++) {
    const step_t *s = &c->steps[i];
    switch (s->op) {
      case OP_CONST: case OP_GET: case OP_GET_GLOBAL: c->fp_is_ext[nf] = 0; c->step_out[i] = (uint32_t)nf++; break;
      case OP_CONST_EXT: c->fp_is_ext[nf] = 1; c->step_out[i] = (uint32_t)nf++; break;
      case OP_ADD: case OP_SUB: case OP_MUL:
        c->fp_is_ext[nf] = c->fp_is_ext[s->a] | c->fp_is_ext[s->b]; c->step_out[i] = (uint32_t)nf++; break;
      case OP_TRUE: c->mix_exp[nm] = 0; c->step_out[i] = (uint32_t)nm++; break;
      case OP_AND_EQZ: c->mix_exp[nm] = c->mix_exp[s->a] + 1; c->step_out[i] = (uint32_t)nm++; break;
      case OP_AND_COND: c->mix_exp[nm] = c->mix_exp[s->a] + c->mix_exp[s->c]; c->step_out[i] = (uint32_t)nm++; break;
      default: fprintf(stderr, "oracle: bad op %u\n", s->op); abort();
    }
  }
  c->n_fp = nf;
  c->n_mix = nm;
  c->max_exp = 0;
  for (size_t i = 0; i < nm; i++)
    if (c->mix_exp[i] > c->max_exp) c->max_exp = c->mix_exp[i];
  c->n_wit_fp = 0;
  for (size_t i = 0; i < c->n_wit_steps; i++) c->n_wit_fp++;
  return c;
}

void or_circuit_free(or_circuit *c) {
  if (!c) return;
  free(c->tap_group); free(c->tap_offset); free(c->tap_back);
  free(c->reg_group); free(c->reg_offset); free(c->reg_combo); free(c->reg_tap_pos); free(c->reg_size);
  free(c->combo_begin); free(c->combo_backs); free(c->steps); free(c->fp_is_ext); free(c->step_out); free(c->mix_exp);
  free(c->wit_steps); free(c->wit_store_col); free(c->wit_store_var); free(c->accum_col); free(c->pin_col);
  free(c->pin_global); free(c->col_kind);
  free(c);
}

/* ---- prover-side evaluation, plain form: one point at a time, every value an fp4 (the definition; the lane-blocked form
 * below must agree with it bit for bit -- tests/test_oracle_circuit.py) ---- */
void or_eval_check_scalar(const or_circuit *c, fp *check, const fp *const groups[3], const fp *globals, const fp *mix,
                          const fp4 *poly_mix, unsigned po2) {
  size_t n = (size_t)1 << po2, domain = 4 * n, mask = domain - 1;
  fp4 *pows = (fp4 *)malloc(sizeof(fp4) * (c->max_exp + 2));
  pows[0] = fp4_one();
  for (size_t i = 1; i <= c->max_exp + 1; i++) pows[i] = fp4_mul(pows[i - 1], *poly_mix);
  /* x^N - 1 takes 4 values: x = 3 w^i, x^N = 3^N * w4^i */
  fp three_n = fp_pow(fp_from_u32(3), n);
  fp w4 = bb_rou_fwd(2);
  fp inv_zero[4];
  for (int k = 0; k < 4; k++) inv_zero[k] = fp_inv(fp_sub(fp_mul(three_n, fp_pow(w4, (uint64_t)k)), FP_ONE));
  const fp *args[2] = {globals, mix};
#pragma omp parallel
  {
    fp4 *fpv = (fp4 *)malloc(sizeof(fp4) * (c->n_fp + 1));
    fp4 *mxv = (fp4 *)malloc(sizeof(fp4) * (c->n_mix + 1));
#pragma omp for schedule(static)
    for (size_t i = 0; i < domain; i++) {
      size_t nf = 0, nm = 0;
      for (size_t k = 0; k < c->n_steps; k++) {
        const step_t *s = &c->steps[k];
        switch (s->op) {
          case OP_CONST: fpv[nf++] = fp4_from_fp(fp_from_u32(s->a)); break;
          case OP_CONST_EXT: {
            fp4 v = {{fp_from_u32(s->a), fp_from_u32(s->b), fp_from_u32(s->c), fp_from_u32(s->d)}};
            fpv[nf++] = v;
          } break;
          case OP_GET: {
            uint32_t t = s->a;
            const fp *g = groups[c->tap_group[t]];
            fpv[nf++] = fp4_from_fp(g[(size_t)c->tap_offset[t] * domain + ((i - 4 * (size_t)c->tap_back[t]) & mask)]);
          } break;
          case OP_GET_GLOBAL: fpv[nf++] = fp4_from_fp(args[s->a][s->b]); break;
          case OP_ADD: fpv[nf] = fp4_add(fpv[s->a], fpv[s->b]); nf++; break;
          case OP_SUB: fpv[nf] = fp4_sub(fpv[s->a], fpv[s->b]); nf++; break;
          case OP_MUL: {
            int ea = c->fp_is_ext[s->a], eb = c->fp_is_ext[s->b];
            if (!ea && !eb) fpv[nf] = fp4_from_fp(fp_mul(fpv[s->a].e[0], fpv[s->b].e[0]));
            else if (ea && !eb) fpv[nf] = fp4_mul_fp(fpv[s->a], fpv[s->b].e[0]);
            else if (!ea && eb) fpv[nf] = fp4_mul_fp(fpv[s->b], fpv[s->a].e[0]);
            else fpv[nf] = fp4_mul(fpv[s->a], fpv[s->b]);
            nf++;
          } break;
          case OP_TRUE: mxv[nm++] = fp4_zero(); break;
          case OP_AND_EQZ: {
            fp4 m = pows[c->mix_exp[s->a]];
            fp4 t = c->fp_is_ext[s->b] ? fp4_mul(m, fpv[s->b]) : fp4_mul_fp(m, fpv[s->b].e[0]);
            mxv[nm] = fp4_add(mxv[s->a], t);
            nm++;
          } break;
          case OP_AND_COND: {
            fp4 m = pows[c->mix_exp[s->a]];
            fp4 t = c->fp_is_ext[s->b] ? fp4_mul(mxv[s->c], fpv[s->b]) : fp4_mul_fp(mxv[s->c], fpv[s->b].e[0]);
            mxv[nm] = fp4_add(mxv[s->a], fp4_mul(t, m));
            nm++;
          } break;
        }
      }
      fp4 r = fp4_mul_fp(mxv[c->ret], inv_zero[i & 3]);
      for (int k = 0; k < 4; k++) check[(size_t)k * domain + i] = r.e[k];
    }
    free(fpv);
    free(mxv);
  }
  free(pows);
}

/* ---- prover-side evaluation, the form the CPU baseline runs: the same program, statically typed (base-field values
 * cost one word and one operation, not four), evaluated on LANES consecutive points at a time so that every field
 * operation is a loop over lanes the compiler vectorises, with value slots recycled at their last use so that the working
 * set stays in cache.  This is the stand-in for upstream's ahead-of-time compiled `poly_fp` (risc0-circuit-rv32im
 * generates straight-line C++ / CUDA from the same step list); exact field arithmetic, so results are identical to
 * or_eval_check_scalar. ---- */
#define LANES 16
enum {
  F_CONST_B, F_CONST_E, F_GET, F_GLOBAL,
  F_ADD_BB, F_ADD_EB, F_ADD_EE, F_SUB_BB, F_SUB_EB, F_SUB_BE, F_SUB_EE, F_MUL_BB, F_MUL_EB, F_MUL_EE,
  F_TRUE, F_EQZ_B, F_EQZ_E, F_COND_B, F_COND_E
};
typedef struct {
  uint32_t op, o, a, b, c; /* slots (base / ext / mix array according to op); F_GET: a = tap; F_GLOBAL: a, b as in the step */
  fp imm[4];               /* constants (Montgomery) */
  uint32_t exp;            /* power of poly_mix of AndEqz / AndCond */
} fstep_t;
typedef struct {
  uint32_t *free_list, n_free, n_slots;
} slot_pool;
static uint32_t slot_take(slot_pool *p) { return p->n_free ? p->free_list[--p->n_free] : p->n_slots++; }
static void slot_give(slot_pool *p, uint32_t s) { p->free_list[p->n_free++] = s; }

typedef fp lane_t[LANES];
static inline void l_add(fp *restrict o, const fp *a, const fp *b) {
  for (int l = 0; l < LANES; l++) o[l] = fp_add(a[l], b[l]);
}
static inline void l_sub(fp *restrict o, const fp *a, const fp *b) {
  for (int l = 0; l < LANES; l++) o[l] = fp_sub(a[l], b[l]);
}
static inline void l_mul(fp *restrict o, const fp *a, const fp *b) {
  for (int l = 0; l < LANES; l++) o[l] = fp_mul(a[l], b[l]);
}
static inline void l_mulc(fp *restrict o, const fp *a, fp k) {
  for (int l = 0; l < LANES; l++) o[l] = fp_mul(a[l], k);
}
static inline void l_copy(fp *restrict o, const fp *a) {
  for (int l = 0; l < LANES; l++) o[l] = a[l];
}
/* ext x ext over lanes: o (4 lane rows) = a * b; o must not alias a or b */
static inline void l_mul_ee(lane_t *restrict o, const lane_t *a, const lane_t *b) {
  for (int l = 0; l < LANES; l++) {
    fp4 x = {{a[0][l], a[1][l], a[2][l], a[3][l]}}, y = {{b[0][l], b[1][l], b[2][l], b[3][l]}};
    fp4 r = fp4_mul(x, y);
    o[0][l] = r.e[0]; o[1][l] = r.e[1]; o[2][l] = r.e[2]; o[3][l] = r.e[3];
  }
}
/* ext lanes x one ext constant */
static inline void l_mul_ec(lane_t *restrict o, const lane_t *a, fp4 k) {
  for (int l = 0; l < LANES; l++) {
    fp4 x = {{a[0][l], a[1][l], a[2][l], a[3][l]}};
    fp4 r = fp4_mul(x, k);
    o[0][l] = r.e[0]; o[1][l] = r.e[1]; o[2][l] = r.e[2]; o[3][l] = r.e[3];
  }
}

void or_eval_check(const or_circuit *c, fp *check, const fp *const groups[3], const fp *globals, const fp *mix,
                   const fp4 *poly_mix, unsigned po2) {
  size_t n = (size_t)1 << po2, domain = 4 * n, mask = domain - 1;
  if (domain < LANES) {
    or_eval_check_scalar(c, check, groups, globals, mix, poly_mix, po2);
    return;
  }
  /* ---- typed program with recycled slots ---- */
  const size_t ns = c->n_steps;
  uint32_t *last_fp = (uint32_t *)malloc((c->n_fp + 1) * 4), *last_mx = (uint32_t *)malloc((c->n_mix + 1) * 4);
  memset(last_fp, 0xff, (c->n_fp + 1) * 4); /* 0xffffffff = never read */
  memset(last_mx, 0xff, (c->n_mix + 1) * 4);
  {
    size_t nf = 0, nm = 0;
    for (size_t k = 0; k < ns; k++) {
      const step_t *s = &c->steps[k];
      switch (s->op) {
        case OP_ADD: case OP_SUB: case OP_MUL: last_fp[s->a] = (uint32_t)k; last_fp[s->b] = (uint32_t)k; nf++; break;
        case OP_TRUE: nm++; break;
        case OP_AND_EQZ: last_mx[s->a] = (uint32_t)k; last_fp[s->b] = (uint32_t)k; nm++; break;
        case OP_AND_COND: last_mx[s->a] = (uint32_t)k; last_fp[s->b] = (uint32_t)k; last_mx[s->c] = (uint32_t)k; nm++; break;
        default: nf++; break;
      }
    }
    (void)nf; (void)nm;
    last_mx[c->ret] = (uint32_t)ns; /* read after the program */
  }
  fstep_t *prog = (fstep_t *)calloc(ns + 1, sizeof(fstep_t));
  uint32_t *slot_fp = (uint32_t *)calloc(c->n_fp + 1, 4), *slot_mx = (uint32_t *)calloc(c->n_mix + 1, 4);
  slot_pool pb = {(uint32_t *)calloc(c->n_fp + 1, 4), 0, 0}, pe = {(uint32_t *)calloc(c->n_fp + 1, 4), 0, 0},
            pm = {(uint32_t *)calloc(c->n_mix + 1, 4), 0, 0};
  {
    size_t nf = 0, nm = 0;
    for (size_t k = 0; k < ns; k++) {
      const step_t *s = &c->steps[k];
      fstep_t *f = &prog[k];
      const int is_fp_out = s->op <= OP_MUL;
      if (is_fp_out) {
        const int ext = c->fp_is_ext[nf];
        f->o = slot_take(ext ? &pe : &pb);
        slot_fp[nf] = f->o;
      } else {
        f->o = slot_take(&pm);
        slot_mx[nm] = f->o;
      }
      switch (s->op) {
        case OP_CONST: f->op = F_CONST_B; f->imm[0] = fp_from_u32(s->a); break;
        case OP_CONST_EXT:
          f->op = F_CONST_E;
          f->imm[0] = fp_from_u32(s->a); f->imm[1] = fp_from_u32(s->b); f->imm[2] = fp_from_u32(s->c); f->imm[3] = fp_from_u32(s->d);
          break;
        case OP_GET: f->op = F_GET; f->a = s->a; break;
        case OP_GET_GLOBAL: f->op = F_GLOBAL; f->a = s->a; f->b = s->b; break;
        case OP_ADD: case OP_SUB: case OP_MUL: {
          const int ea = c->fp_is_ext[s->a], eb = c->fp_is_ext[s->b];
          f->a = slot_fp[s->a]; f->b = slot_fp[s->b];
          if (s->op == OP_SUB) {
            f->op = !ea && !eb ? F_SUB_BB : ea && !eb ? F_SUB_EB : !ea && eb ? F_SUB_BE : F_SUB_EE;
          } else {
            const uint32_t bb = s->op == OP_ADD ? F_ADD_BB : F_MUL_BB, eb_ = s->op == OP_ADD ? F_ADD_EB : F_MUL_EB,
                           ee = s->op == OP_ADD ? F_ADD_EE : F_MUL_EE;
            if (!ea && !eb) f->op = bb;
            else if (ea && eb) f->op = ee;
            else {
              f->op = eb_;            /* commutative: put the ext operand first */
              if (!ea) { uint32_t t = f->a; f->a = f->b; f->b = t; }
            }
          }
        } break;
        case OP_TRUE: f->op = F_TRUE; break;
        case OP_AND_EQZ:
          f->op = c->fp_is_ext[s->b] ? F_EQZ_E : F_EQZ_B;
          f->a = slot_mx[s->a]; f->b = slot_fp[s->b]; f->exp = c->mix_exp[s->a];
          break;
        case OP_AND_COND:
          f->op = c->fp_is_ext[s->b] ? F_COND_E : F_COND_B;
          f->a = slot_mx[s->a]; f->b = slot_fp[s->b]; f->c = slot_mx[s->c]; f->exp = c->mix_exp[s->a];
          break;
      }
      /* operands that die here give their slots back (after the output slot was taken: no aliasing inside a step) */
      if (s->op == OP_ADD || s->op == OP_SUB || s->op == OP_MUL) {
        if (last_fp[s->a] == k) slot_give(c->fp_is_ext[s->a] ? &pe : &pb, slot_fp[s->a]);
        if (s->b != s->a && last_fp[s->b] == k) slot_give(c->fp_is_ext[s->b] ? &pe : &pb, slot_fp[s->b]);
      } else if (s->op == OP_AND_EQZ || s->op == OP_AND_COND) {
        if (last_mx[s->a] == k) slot_give(&pm, slot_mx[s->a]);
        if (last_fp[s->b] == k) slot_give(c->fp_is_ext[s->b] ? &pe : &pb, slot_fp[s->b]);
        if (s->op == OP_AND_COND && s->c != s->a && last_mx[s->c] == k) slot_give(&pm, slot_mx[s->c]);
      }
      /* a value nobody reads gives its slot back at once */
      if (is_fp_out) {
        if (last_fp[nf] == 0xffffffffu) slot_give(c->fp_is_ext[nf] ? &pe : &pb, f->o);
        nf++;
      } else {
        if (last_mx[nm] == 0xffffffffu) slot_give(&pm, f->o);
        nm++;
      }
    }
  }
  const uint32_t ret_slot = slot_mx[c->ret];
  const size_t nb_slots = pb.n_slots + 1, ne_slots = pe.n_slots + 1, nm_slots = pm.n_slots + 1;
  free(last_fp); free(last_mx); free(slot_fp); free(slot_mx); free(pb.free_list); free(pe.free_list); free(pm.free_list);

  fp4 *pows = (fp4 *)malloc(sizeof(fp4) * (c->max_exp + 2));
  pows[0] = fp4_one();
  for (size_t i = 1; i <= c->max_exp + 1; i++) pows[i] = fp4_mul(pows[i - 1], *poly_mix);
  fp three_n = fp_pow(fp_from_u32(3), n);
  fp w4 = bb_rou_fwd(2);
  fp inv_zero[4];
  for (int k = 0; k < 4; k++) inv_zero[k] = fp_inv(fp_sub(fp_mul(three_n, fp_pow(w4, (uint64_t)k)), FP_ONE));
  const fp *args[2] = {globals, mix};
#pragma omp parallel
  {
    lane_t *B = (lane_t *)aligned_alloc(64, sizeof(lane_t) * nb_slots);
    lane_t *E = (lane_t *)aligned_alloc(64, sizeof(lane_t) * 4 * ne_slots);
    lane_t *M = (lane_t *)aligned_alloc(64, sizeof(lane_t) * 4 * nm_slots);
    lane_t T[4], U[4];
#pragma omp for schedule(static)
    for (size_t i = 0; i < domain; i += LANES) {
      for (size_t k = 0; k < ns; k++) {
        const fstep_t *f = &prog[k];
        switch (f->op) {
          case F_CONST_B: for (int l = 0; l < LANES; l++) B[f->o][l] = f->imm[0]; break;
          case F_CONST_E:
            for (int q = 0; q < 4; q++)
              for (int l = 0; l < LANES; l++) E[4 * f->o + q][l] = f->imm[q];
            break;
          case F_GET: {
            const uint32_t t = f->a;
            const fp *col = groups[c->tap_group[t]] + (size_t)c->tap_offset[t] * domain;
            const size_t j0 = (i - 4 * (size_t)c->tap_back[t]) & mask;
            if (j0 + LANES <= domain) l_copy(B[f->o], col + j0);
            else
              for (int l = 0; l < LANES; l++) B[f->o][l] = col[(j0 + (size_t)l) & mask];
          } break;
          case F_GLOBAL: { const fp g = args[f->a][f->b]; for (int l = 0; l < LANES; l++) B[f->o][l] = g; } break;
          case F_ADD_BB: l_add(B[f->o], B[f->a], B[f->b]); break;
          case F_SUB_BB: l_sub(B[f->o], B[f->a], B[f->b]); break;
          case F_MUL_BB: l_mul(B[f->o], B[f->a], B[f->b]); break;
          case F_ADD_EB: /* ext a + base b */
            l_add(E[4 * f->o], E[4 * f->a], B[f->b]);
            for (int q = 1; q < 4; q++) l_copy(E[4 * f->o + q], E[4 * f->a + q]);
            break;
          case F_SUB_EB: /* ext a - base b */
            l_sub(E[4 * f->o], E[4 * f->a], B[f->b]);
            for (int q = 1; q < 4; q++) l_copy(E[4 * f->o + q], E[4 * f->a + q]);
            break;
          case F_SUB_BE: /* base a - ext b */
            l_sub(E[4 * f->o], B[f->a], E[4 * f->b]);
            for (int q = 1; q < 4; q++)
              for (int l = 0; l < LANES; l++) E[4 * f->o + q][l] = fp_neg(E[4 * f->b + q][l]);
            break;
          case F_ADD_EE: for (int q = 0; q < 4; q++) l_add(E[4 * f->o + q], E[4 * f->a + q], E[4 * f->b + q]); break;
          case F_SUB_EE: for (int q = 0; q < 4; q++) l_sub(E[4 * f->o + q], E[4 * f->a + q], E[4 * f->b + q]); break;
          case F_MUL_EB: for (int q = 0; q < 4; q++) l_mul(E[4 * f->o + q], E[4 * f->a + q], B[f->b]); break;
          case F_MUL_EE: l_mul_ee(&E[4 * f->o], (const lane_t *)&E[4 * f->a], (const lane_t *)&E[4 * f->b]); break;
          case F_TRUE:
            for (int q = 0; q < 4; q++)
              for (int l = 0; l < LANES; l++) M[4 * f->o + q][l] = 0;
            break;
          case F_EQZ_B: { /* tot + pow * val */
            const fp4 m = pows[f->exp];
            for (int q = 0; q < 4; q++) {
              l_mulc(T[q], B[f->b], m.e[q]);
              l_add(M[4 * f->o + q], M[4 * f->a + q], T[q]);
            }
          } break;
          case F_EQZ_E: {
            l_mul_ec(T, (const lane_t *)&E[4 * f->b], pows[f->exp]);
            for (int q = 0; q < 4; q++) l_add(M[4 * f->o + q], M[4 * f->a + q], T[q]);
          } break;
          case F_COND_B: { /* tot + (inner * cond) * pow */
            for (int q = 0; q < 4; q++) l_mul(T[q], M[4 * f->c + q], B[f->b]);
            l_mul_ec(U, (const lane_t *)T, pows[f->exp]);
            for (int q = 0; q < 4; q++) l_add(M[4 * f->o + q], M[4 * f->a + q], U[q]);
          } break;
          case F_COND_E: {
            l_mul_ee(T, (const lane_t *)&M[4 * f->c], (const lane_t *)&E[4 * f->b]);
            l_mul_ec(U, (const lane_t *)T, pows[f->exp]);
            for (int q = 0; q < 4; q++) l_add(M[4 * f->o + q], M[4 * f->a + q], U[q]);
          } break;
        }
      }
      for (int q = 0; q < 4; q++)
        for (int l = 0; l < LANES; l++) check[(size_t)q * domain + i + (size_t)l] = fp_mul(M[4 * ret_slot + q][l], inv_zero[l & 3]);
    }
    free(B); free(E); free(M);
  }
  free(pows);
  free(prog);
}

/* adapter.rs poly_ext: same program on extension-field tap evaluations; MixState{tot, mul} carried explicitly. */
fp4 or_poly_ext(const or_circuit *c, const fp4 *poly_mix, const fp4 *eval_u, const fp *globals, const fp *mix) {
  fp4 *fpv = (fp4 *)malloc(sizeof(fp4) * (c->n_fp + 1));
  fp4 *tot = (fp4 *)malloc(sizeof(fp4) * (c->n_mix + 1));
  fp4 *mul = (fp4 *)malloc(sizeof(fp4) * (c->n_mix + 1));
  const fp *args[2] = {globals, mix};
  size_t nf = 0, nm = 0;
  for (size_t k = 0; k < c->n_steps; k++) {
    const step_t *s = &c->steps[k];
    switch (s->op) {
      case OP_CONST: fpv[nf++] = fp4_from_fp(fp_from_u32(s->a)); break;
      case OP_CONST_EXT: {
        fp4 v = {{fp_from_u32(s->a), fp_from_u32(s->b), fp_from_u32(s->c), fp_from_u32(s->d)}};
        fpv[nf++] = v;
      } break;
      case OP_GET: fpv[nf++] = eval_u[s->a]; break;
      case OP_GET_GLOBAL: fpv[nf++] = fp4_from_fp(args[s->a][s->b]); break;
      case OP_ADD: fpv[nf] = fp4_add(fpv[s->a], fpv[s->b]); nf++; break;
      case OP_SUB: fpv[nf] = fp4_sub(fpv[s->a], fpv[s->b]); nf++; break;
      case OP_MUL: fpv[nf] = fp4_mul(fpv[s->a], fpv[s->b]); nf++; break;
      case OP_TRUE: tot[nm] = fp4_zero(); mul[nm] = fp4_one(); nm++; break;
      case OP_AND_EQZ:
        tot[nm] = fp4_add(tot[s->a], fp4_mul(mul[s->a], fpv[s->b]));
        mul[nm] = fp4_mul(mul[s->a], *poly_mix);
        nm++;
        break;
      case OP_AND_COND:
        tot[nm] = fp4_add(tot[s->a], fp4_mul(fp4_mul(fpv[s->b], tot[s->c]), mul[s->a]));
        mul[nm] = fp4_mul(mul[s->a], mul[s->c]);
        nm++;
        break;
    }
  }
  fp4 r = tot[c->ret];
  free(fpv); free(tot); free(mul);
  return r;
}

/* ---- synthetic witness ---- */
static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
/* counter-based value stream shared (by specification, see DESIGN.md "synthetic witness") with the CUDA witgen */
uint32_t or_synth_rand(uint64_t seed, uint32_t stream, uint32_t col, uint64_t row) {
  uint64_t z = splitmix64(seed ^ splitmix64(((uint64_t)stream << 56) ^ ((uint64_t)col << 32) ^ row));
  return (uint32_t)(z % BB_P);
}
enum { STREAM_CODE = 1, STREAM_FREE = 2, STREAM_NOISE = 3, STREAM_ACCUM_NOISE = 4 };

void or_witgen(const or_circuit *c, unsigned po2, fp *code, fp *data, fp *globals, uint64_t seed, size_t zk_rows) {
  size_t n = (size_t)1 << po2, mask = n - 1, active_rows = n - zk_rows;
  size_t wc = c->group_size[1], wd = c->group_size[2];
  /* code group: col 0 = active selector, col 1 = first-row selector, rest = per-po2 deterministic constants */
#pragma omp parallel for schedule(static)
  for (size_t col = 0; col < wc; col++)
    for (size_t r = 0; r < n; r++) {
      fp v;
      if (col == 0) v = r < active_rows ? FP_ONE : 0;
      else if (col == 1) v = r == 0 ? FP_ONE : 0;
      else v = or_synth_rand(0xC0DE0000ull + po2, STREAM_CODE, (uint32_t)col, r);
      code[col * n + r] = v;
    }
  /* free columns; noise rows of every column */
#pragma omp parallel for schedule(static)
  for (size_t col = 0; col < wd; col++)
    for (size_t r = 0; r < n; r++) {
      fp v = 0;
      if (r >= active_rows) v = or_synth_rand(seed, STREAM_NOISE, (uint32_t)col, r);
      else if (c->col_kind[col] == 0) v = or_synth_rand(seed, STREAM_FREE, (uint32_t)col, r);
      else if (c->col_kind[col] == 1) v = (or_synth_rand(seed, STREAM_FREE, (uint32_t)col, r) & 1) ? FP_ONE : 0;
      else if ((c->col_kind[col] & 0xff) == 3) { /* one-hot group: member idx is 1 on the rows whose group draw selects it */
        uint32_t k = c->col_kind[col], idx = (k >> 8) & 0xff, size = (k >> 16) & 0xff, grp = k >> 24;
        v = (or_synth_rand(seed, STREAM_FREE, 0x8000u + grp, r) % size) == idx ? FP_ONE : 0;
      }
      data[col * n + r] = v;
    }
  for (size_t g = 0; g < c->globals_size; g++) globals[g] = or_synth_rand(seed, STREAM_FREE, 0xffff, g);
  /* claim words (protocol_params.h, synthetic sessions): post-state of seed s == pre-state of seed s + 1 */
  if (pp_claim_post_offset(c->globals_size) >= 0)
    for (size_t g = (size_t)pp_claim_post_offset(c->globals_size); g < c->globals_size; g++)
      globals[g] = or_synth_rand(seed + 1, STREAM_FREE, 0xffff, g - PP_CLAIM_DIGEST_WORDS);
  for (size_t i = 0; i < c->n_pins; i++) globals[c->pin_global[i]] = data[(size_t)c->pin_col[i] * n];
  /* derived columns: run the witgen program on every active row (taps read the trace domain, stride 1) */
  const fp *grp[3] = {NULL, code, data};
#pragma omp parallel
  {
    fp *v = (fp *)malloc(sizeof(fp) * (c->n_wit_steps + 1));
#pragma omp for schedule(static)
    for (size_t r = 0; r < active_rows; r++) {
      size_t nf = 0;
      for (size_t k = 0; k < c->n_wit_steps; k++) {
        const step_t *s = &c->wit_steps[k];
        switch (s->op) {
          case OP_CONST: v[nf++] = fp_from_u32(s->a); break;
          case OP_GET: {
            uint32_t t = s->a;
            v[nf++] = grp[c->tap_group[t]][(size_t)c->tap_offset[t] * n + ((r - c->tap_back[t]) & mask)];
          } break;
          case OP_GET_GLOBAL: v[nf++] = globals[s->b]; break;
          case OP_ADD: v[nf] = fp_add(v[s->a], v[s->b]); nf++; break;
          case OP_SUB: v[nf] = fp_sub(v[s->a], v[s->b]); nf++; break;
          case OP_MUL: v[nf] = fp_mul(v[s->a], v[s->b]); nf++; break;
          default: abort();
        }
      }
      for (size_t k = 0; k < c->n_wit_stores; k++) data[(size_t)c->wit_store_col[k] * n + r] = v[c->wit_store_var[k]];
    }
    free(v);
  }
}

/* accumulate: accumulator k is the ext running sum over active rows of 1/(M_k + data[accum_col[k]][r]),
 * M_k = ext(mix[4k..4k+3]); the last zk_rows rows take the supplied noise ([4*n_accum][zk_rows]). */
void or_accumulate(const or_circuit *c, unsigned po2, fp *accum, const fp *code, const fp *data, const fp *mix,
                   const fp *accum_noise, size_t zk_rows) {
  (void)code;
  size_t n = (size_t)1 << po2, active_rows = n - zk_rows;
#pragma omp parallel for schedule(dynamic)
  for (size_t k = 0; k < c->n_accum; k++) {
    fp4 m = {{mix[4 * k], mix[4 * k + 1], mix[4 * k + 2], mix[4 * k + 3]}};
    const fp *d = data + (size_t)c->accum_col[k] * n;
    fp4 run = fp4_zero();
    for (size_t r = 0; r < active_rows; r++) {
      fp4 t = m;
      t.e[0] = fp_add(t.e[0], d[r]);
      run = fp4_add(run, fp4_inv(t));
      for (int j = 0; j < 4; j++) accum[(4 * k + j) * n + r] = run.e[j];
    }
    for (int j = 0; j < 4; j++)
      for (size_t r = active_rows; r < n; r++) accum[(4 * k + j) * n + r] = accum_noise[(4 * k + j) * zk_rows + (r - active_rows)];
  }
}
