/* ORACLE (test infrastructure, not product code): the `Hal` operations of the segment prover on host memory.
 *
 * Restates risc0-zkp 3.0.2 `hal/cpu.rs` (CpuHal) and `core/ntt.rs` (un-vendored dependency of /root/reference,
 * Cargo.lock:5393-5396) following SURVEY.md Appendix A.3/A.5/A.6/A.10 and the scope table §8(a) rows a5-a16.
 * All matrices are column-major: element (col c, row r) lives at c*rows + r.  Values are Montgomery words < P.
 * PARITY: "parity unpinned" against the reference (no golden vectors exist for these ops); pinned instead by
 * algebraic identities in tests/test_oracle_hal.py (naive DFT, Horner, fold identity, Merkle re-hash).
 */
#include "oracle.h"
#include <omp.h>
#include <stdlib.h>
#include <string.h>

int or_num_threads(void) { return omp_get_max_threads(); }
/* the OpenMP runtime reads OMP_NUM_THREADS when it is first loaded (torch loads it long before the oracle): set it explicitly */
void or_set_num_threads(int n) { if (n > 0) omp_set_num_threads(n); }
void or_free(void *p) { free(p); }

/* twiddle table w^0..w^(n/2-1) */
static fp *make_twiddles(fp w, size_t half) {
  fp *tw = (fp *)malloc(sizeof(fp) * (half ? half : 1));
  fp cur = FP_ONE;
  for (size_t i = 0; i < half; i++) {
    tw[i] = cur;
    cur = fp_mul(cur, w);
  }
  return tw;
}

/* core/ntt.rs interpolate_ntt: natural-order evaluations -> bit-reversed coefficients, scaled by 1/n
 * (decimation in frequency with ROU_REV twiddles). */
static void interpolate_one(fp *io, size_t n, const fp *tw, fp norm) {
  unsigned bits = bb_log2(n);
  for (unsigned s = bits; s >= 1; s--) {
    size_t half = (size_t)1 << (s - 1);
    size_t tstep = n >> s;
    for (size_t blk = 0; blk < n; blk += 2 * half) {
      fp *p = io + blk;
      for (size_t i = 0; i < half; i++) {
        fp a = p[i], b = p[i + half];
        p[i] = fp_add(a, b);
        p[i + half] = fp_mul(fp_sub(a, b), tw[i * tstep]);
      }
    }
  }
  for (size_t i = 0; i < n; i++) io[i] = fp_mul(io[i], norm);
}

/* core/ntt.rs evaluate_ntt: bit-reversed coefficients -> natural-order evaluations, skipping the first
 * `expand_bits` butterfly layers (decimation in time with ROU_FWD twiddles). */
static void evaluate_one(fp *io, size_t n, const fp *tw, unsigned expand_bits) {
  unsigned bits = bb_log2(n);
  for (unsigned s = expand_bits + 1; s <= bits; s++) {
    size_t half = (size_t)1 << (s - 1);
    size_t tstep = n >> s;
    for (size_t blk = 0; blk < n; blk += 2 * half) {
      fp *p = io + blk;
      for (size_t i = 0; i < half; i++) {
        fp a = p[i], b = fp_mul(p[i + half], tw[i * tstep]);
        p[i] = fp_add(a, b);
        p[i + half] = fp_sub(a, b);
      }
    }
  }
}

/* Few columns (the 4-column check / FRI groups): parallel over the butterflies of each layer instead of over columns, so
 * that a 4-column transform still uses every host thread.  Same butterflies, same results. */
static void interpolate_layers_par(fp *io, size_t n, const fp *tw, fp norm) {
  unsigned bits = bb_log2(n);
  for (unsigned s = bits; s >= 1; s--) {
    const size_t half = (size_t)1 << (s - 1), tstep = n >> s;
#pragma omp parallel for schedule(static)
    for (size_t j = 0; j < n / 2; j++) {
      const size_t i = j & (half - 1);
      fp *p = io + ((j >> (s - 1)) << s);
      fp a = p[i], b = p[i + half];
      p[i] = fp_add(a, b);
      p[i + half] = fp_mul(fp_sub(a, b), tw[i * tstep]);
    }
  }
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) io[i] = fp_mul(io[i], norm);
}
static void evaluate_layers_par(fp *io, size_t n, const fp *tw, unsigned expand_bits) {
  unsigned bits = bb_log2(n);
  for (unsigned s = expand_bits + 1; s <= bits; s++) {
    const size_t half = (size_t)1 << (s - 1), tstep = n >> s;
#pragma omp parallel for schedule(static)
    for (size_t j = 0; j < n / 2; j++) {
      const size_t i = j & (half - 1);
      fp *p = io + ((j >> (s - 1)) << s);
      fp a = p[i], b = fp_mul(p[i + half], tw[i * tstep]);
      p[i] = fp_add(a, b);
      p[i + half] = fp_sub(a, b);
    }
  }
}
static int few_columns(size_t count, size_t n) { return count < (size_t)omp_get_max_threads() && n >= 4096; }

void or_batch_interpolate_ntt(fp *io, size_t count, size_t n) {
  unsigned bits = bb_log2(n);
  fp *tw = make_twiddles(bb_rou_rev(bits), n / 2);
  fp norm = fp_inv(fp_from_u32((uint32_t)n));
  if (few_columns(count, n)) {
    for (size_t c = 0; c < count; c++) interpolate_layers_par(io + c * n, n, tw, norm);
  } else {
#pragma omp parallel for schedule(dynamic)
    for (size_t c = 0; c < count; c++) interpolate_one(io + c * n, n, tw, norm);
  }
  free(tw);
}

void or_batch_evaluate_ntt(fp *io, size_t count, size_t n, unsigned expand_bits) {
  unsigned bits = bb_log2(n);
  fp *tw = make_twiddles(bb_rou_fwd(bits), n / 2);
  if (few_columns(count, n)) {
    for (size_t c = 0; c < count; c++) evaluate_layers_par(io + c * n, n, tw, expand_bits);
  } else {
#pragma omp parallel for schedule(dynamic)
    for (size_t c = 0; c < count; c++) evaluate_one(io + c * n, n, tw, expand_bits);
  }
  free(tw);
}

/* hal/cpu.rs batch_expand_into_evaluate_ntt: out[i] = in[i >> expand_bits], then the remaining layers. */
void or_batch_expand_into_evaluate_ntt(fp *out, const fp *in, size_t count, size_t in_n, unsigned expand_bits) {
  size_t out_n = in_n << expand_bits;
  unsigned bits = bb_log2(out_n);
  fp *tw = make_twiddles(bb_rou_fwd(bits), out_n / 2);
  if (few_columns(count, out_n)) {
    for (size_t c = 0; c < count; c++) {
      fp *o = out + c * out_n;
      const fp *p = in + c * in_n;
#pragma omp parallel for schedule(static)
      for (size_t i = 0; i < out_n; i++) o[i] = p[i >> expand_bits];
      evaluate_layers_par(o, out_n, tw, expand_bits);
    }
  } else {
#pragma omp parallel for schedule(dynamic)
    for (size_t c = 0; c < count; c++) {
      fp *o = out + c * out_n;
      const fp *p = in + c * in_n;
      for (size_t i = 0; i < out_n; i++) o[i] = p[i >> expand_bits];
      evaluate_one(o, out_n, tw, expand_bits);
    }
  }
  free(tw);
}

void or_batch_bit_reverse(fp *io, size_t count, size_t n) {
  unsigned bits = bb_log2(n);
#pragma omp parallel for schedule(dynamic)
  for (size_t c = 0; c < count; c++) {
    fp *p = io + c * n;
    for (size_t i = 0; i < n; i++) {
      size_t r = bb_bit_rev((uint32_t)i, bits);
      if (i < r) {
        fp t = p[i];
        p[i] = p[r];
        p[r] = t;
      }
    }
  }
}

/* hal/cpu.rs zk_shift: the coefficient in bit-reversed slot j is multiplied by 3^rev(j)  (f(x) -> f(3x)). */
void or_zk_shift(fp *io, size_t count, size_t n) {
  unsigned bits = bb_log2(n);
  fp three = fp_from_u32(3);
  fp *pw = (fp *)malloc(sizeof(fp) * n); /* pw[j] = 3^rev(j) */
  fp cur = FP_ONE;
  for (size_t i = 0; i < n; i++) {
    pw[bb_bit_rev((uint32_t)i, bits)] = cur;
    cur = fp_mul(cur, three);
  }
#pragma omp parallel for schedule(static)
  for (size_t c = 0; c < count; c++) {
    fp *p = io + c * n;
    for (size_t j = 0; j < n; j++) p[j] = fp_mul(p[j], pw[j]);
  }
  free(pw);
}

/* hal/cpu.rs hash_rows: digest r = unpadded_hash(matrix[0][r], matrix[1][r], ...). */
void or_hash_rows_scalar(uint32_t *out, const fp *matrix, size_t rows, size_t cols) {
  or_poseidon2_init();
#pragma omp parallel for schedule(static)
  for (size_t r = 0; r < rows; r++) or_hash_elems_strided(matrix + r, cols, rows, out + 8 * r);
}
/* 16 consecutive rows per sponge (the matrix is column-major: 16 rows of a column are one contiguous load) */
void or_hash_rows(uint32_t *out, const fp *matrix, size_t rows, size_t cols) {
  or_poseidon2_init();
  const size_t blocks = rows / P2_LANES;
#pragma omp parallel for schedule(static)
  for (size_t b = 0; b < blocks; b++) {
    fp st[24][P2_LANES];
    memset(st, 0, sizeof st);
    const size_t r0 = b * P2_LANES;
    size_t unmixed = 0;
    for (size_t c = 0; c < cols; c++) {
      memcpy(st[unmixed++], matrix + c * rows + r0, sizeof(fp) * P2_LANES);
      if (unmixed == 16) {
        or_poseidon2_mix_lanes(st);
        unmixed = 0;
      }
    }
    if (unmixed != 0 || cols == 0) {
      for (size_t i = unmixed; i < 16; i++) memset(st[i], 0, sizeof(fp) * P2_LANES);
      or_poseidon2_mix_lanes(st);
    }
    for (int l = 0; l < P2_LANES; l++)
      for (int k = 0; k < 8; k++) out[8 * (r0 + (size_t)l) + k] = st[k][l];
  }
  for (size_t r = blocks * P2_LANES; r < rows; r++) or_hash_elems_strided(matrix + r, cols, rows, out + 8 * r);
}

/* hal/cpu.rs hash_fold: node[out_size + j] = hash_pair(node[in_size + 2j], node[in_size + 2j + 1]) */
void or_hash_fold_scalar(uint32_t *nodes, size_t in_size, size_t out_size) {
  or_poseidon2_init();
#pragma omp parallel for schedule(static)
  for (size_t j = 0; j < out_size; j++)
    or_hash_pair(nodes + 8 * (in_size + 2 * j), nodes + 8 * (in_size + 2 * j + 1), nodes + 8 * (out_size + j));
}
void or_hash_fold(uint32_t *nodes, size_t in_size, size_t out_size) {
  or_poseidon2_init();
  const size_t blocks = out_size / P2_LANES;
#pragma omp parallel for schedule(static)
  for (size_t b = 0; b < blocks; b++) {
    fp st[24][P2_LANES];
    for (int l = 0; l < P2_LANES; l++) {
      const uint32_t *pair = nodes + 8 * (in_size + 2 * (b * P2_LANES + (size_t)l)); /* 16 consecutive words: left ‖ right */
      for (int k = 0; k < 16; k++) st[k][l] = pair[k];
      for (int k = 16; k < 24; k++) st[k][l] = 0;
    }
    or_poseidon2_mix_lanes(st);
    for (int l = 0; l < P2_LANES; l++)
      for (int k = 0; k < 8; k++) nodes[8 * (out_size + b * P2_LANES + (size_t)l) + k] = st[k][l];
  }
  for (size_t j = blocks * P2_LANES; j < out_size; j++)
    or_hash_pair(nodes + 8 * (in_size + 2 * j), nodes + 8 * (in_size + 2 * j + 1), nodes + 8 * (out_size + j));
}

/* hal/cpu.rs fri_fold: out[idx] = sum_{i<16} mix^i * in[rev4(i)*count + idx] on [4][.] component-major buffers. */
void or_fri_fold(fp *out, const fp *in, const fp4 *mix, size_t count) {
  size_t in_count = count * OR_FRI_FOLD;
#pragma omp parallel for schedule(static)
  for (size_t idx = 0; idx < count; idx++) {
    fp4 tot = fp4_zero(), cur = fp4_one();
    for (unsigned i = 0; i < OR_FRI_FOLD; i++) {
      size_t rev_idx = (size_t)bb_bit_rev(i, 4) * count + idx;
      fp4 f = {{in[rev_idx], in[in_count + rev_idx], in[2 * in_count + rev_idx], in[3 * in_count + rev_idx]}};
      tot = fp4_add(tot, fp4_mul(cur, f));
      cur = fp4_mul(cur, *mix);
    }
    for (int k = 0; k < 4; k++) out[k * count + idx] = tot.e[k];
  }
}

/* hal/cpu.rs batch_evaluate_any: out[i] = sum_j coeffs[which[i]*n + j] * xs[i]^j (natural-order coefficients). */
void or_batch_evaluate_any(const fp *coeffs, size_t poly_count, size_t n, const uint32_t *which, const fp4 *xs,
                           size_t eval_count, fp4 *out) {
  (void)poly_count;
#pragma omp parallel for schedule(dynamic)
  for (size_t i = 0; i < eval_count; i++) {
    const fp *p = coeffs + (size_t)which[i] * n;
    fp4 x = xs[i], tot = fp4_zero();
    for (size_t j = n; j-- > 0;) tot = fp4_add(fp4_mul(tot, x), fp4_from_fp(p[j]));
    out[i] = tot;
  }
}

/* hal/cpu.rs mix_poly_coeffs: out[combos[i]*count + idx] += mix_start*mix^i * in[i*count + idx]. */
void or_mix_poly_coeffs(fp4 *out, const fp4 *mix_start, const fp4 *mix, const fp *in, const uint32_t *combos,
                        size_t input_size, size_t count) {
#pragma omp parallel for schedule(static)
  for (size_t idx = 0; idx < count; idx++) {
    fp4 cur = *mix_start;
    for (size_t i = 0; i < input_size; i++) {
      size_t o = (size_t)combos[i] * count + idx;
      out[o] = fp4_add(out[o], fp4_mul_fp(cur, in[i * count + idx]));
      cur = fp4_mul(cur, *mix);
    }
  }
}

void or_eltwise_add_elem(fp *out, const fp *a, const fp *b, size_t n) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++) out[i] = fp_add(a[i], b[i]);
}
/* hal/cpu.rs eltwise_sum_extelem: AoS ext input [to_add][count] -> component-major output [4][count]. */
void or_eltwise_sum_extelem(fp *out, const fp4 *in, size_t count, size_t to_add) {
#pragma omp parallel for schedule(static)
  for (size_t idx = 0; idx < count; idx++) {
    fp4 tot = fp4_zero();
    for (size_t i = 0; i < to_add; i++) tot = fp4_add(tot, in[i * count + idx]);
    for (int k = 0; k < 4; k++) out[k * count + idx] = tot.e[k];
  }
}
void or_eltwise_copy_elem(fp *out, const fp *in, size_t n) { memcpy(out, in, n * sizeof(fp)); }
/* Elem::INVALID (0xffffffff) marks unset trace cells; zeroize maps it to 0 (hal/cpu.rs eltwise_zeroize_elem). */
void or_eltwise_zeroize_elem(fp *io, size_t n) {
#pragma omp parallel for schedule(static)
  for (size_t i = 0; i < n; i++)
    if (io[i] == BB_INVALID) io[i] = 0;
}
/* hal/cpu.rs gather_sample: dst[i] = src[idx + i*stride] (a Merkle row out of a column-major matrix). */
void or_gather_sample(fp *dst, const fp *src, size_t idx, size_t size, size_t stride) {
  for (size_t i = 0; i < size; i++) dst[i] = src[idx + i * stride];
}

void or_fp_mul_vec(fp *out, const fp *a, const fp *b, size_t n) {
  for (size_t i = 0; i < n; i++) out[i] = fp_mul(a[i], b[i]);
}
void or_fp4_mul_vec(fp4 *out, const fp4 *a, const fp4 *b, size_t n) {
  for (size_t i = 0; i < n; i++) out[i] = fp4_mul(a[i], b[i]);
}
void or_fp4_inv_vec(fp4 *out, const fp4 *a, size_t n) {
  for (size_t i = 0; i < n; i++) out[i] = fp4_inv(a[i]);
}
void or_to_montgomery(uint32_t *io, size_t n) {
  for (size_t i = 0; i < n; i++) io[i] = fp_from_u32(io[i]);
}
void or_from_montgomery(uint32_t *io, size_t n) {
  for (size_t i = 0; i < n; i++) io[i] = fp_to_u32(io[i]);
}

/* hal/cpu.rs scatter [RECALL lo-med]: CSR form, cycle c writes into[offsets[i]] = values[i] for i in [index[c], index[c+1]) */
void or_scatter(fp *into, const uint32_t *index, const uint32_t *offsets, const fp *values, size_t count) {
  for (size_t c = 0; c < count; c++)
    for (uint32_t i = index[c]; i < index[c + 1]; i++) into[offsets[i]] = values[i];
}
/* hal/cpu.rs prefix_products [RECALL med]: io[i] = io[0] * ... * io[i] over AoS ext elements */
void or_prefix_products(fp4 *io, size_t n) {
  for (size_t i = 1; i < n; i++) io[i] = fp4_mul(io[i - 1], io[i]);
}
