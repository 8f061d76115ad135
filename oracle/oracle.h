/* ORACLE — test infrastructure, NOT product code.
 *
 * A plain-C CPU restatement of the RISC Zero STARK segment-prover hot path that zeth reaches through
 * `default_prover().prove(env, elf)` (/root/reference crates/host/src/lib.rs:137).  The arithmetic lives in
 * un-vendored crates.io dependencies (risc0-zkp 3.0.2, risc0-core 3.0.0, risc0-circuit-rv32im 4.0.2;
 * Cargo.lock:5393,5338,5320) whose sources are absent from /root/reference, so every function here restates
 * the published algorithm from SURVEY.md Appendix A and cites the upstream file it follows.
 *
 * PARITY STATUS: field + Poseidon2 permutation are pinned by known-answer vectors;
 * everything above them (sponge framing, Merkle, FRI, transcript order, seal layout) is **parity unpinned**
 * (no golden vector of the reference exists in this environment) and is anchored by algebraic identities and
 * by the verifier restatement in zkp.c accepting the prover's seals.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use this library.
 */
#ifndef ORACLE_H
#define ORACLE_H
#include "bb.h"

#ifdef __cplusplus
extern "C" {
#endif

/* protocol constants: risc0-zkp 3.0.2 src/lib.rs.  The recalled conventions live in ONE header shared
 * with the product (values + confidence tags): a real risc0 seal flips them in one place for oracle, prover, verifier. */
#include "../include/protocol_params.h"
#define OR_INV_RATE PP_INV_RATE
#define OR_QUERIES PP_QUERIES
#define OR_FRI_FOLD PP_FRI_FOLD
#define OR_FRI_MIN_DEGREE PP_FRI_MIN_DEGREE
#define OR_EXT_SIZE PP_EXT_SIZE
#define OR_CHECK_SIZE PP_CHECK_SIZE
#define OR_DIGEST_WORDS PP_DIGEST_WORDS
#define OR_ZK_CYCLES PP_ZK_CYCLES

/* ---- Poseidon2 (poseidon2.c) ---- */
typedef struct {
  fp cells[24];
  unsigned pool_used;
} or_rng;
void or_poseidon2_init(void);
void or_poseidon2_mix(fp *state24);
#define P2_LANES 16
void or_poseidon2_mix_lanes(fp (*state)[P2_LANES]); /* state[24][P2_LANES]: the same permutation on 16 states at once */
void or_hash_elems_strided(const fp *in, size_t n, size_t stride, uint32_t *digest);
void or_hash_elem_slice(const fp *in, size_t n, uint32_t *digest);
void or_hash_ext_elem_slice(const fp4 *in, size_t n, uint32_t *digest);
void or_hash_pair(const uint32_t *a, const uint32_t *b, uint32_t *digest);
void or_rng_init(or_rng *r);
void or_rng_mix(or_rng *r, const uint32_t *digest);
fp or_rng_elem(or_rng *r);
fp4 or_rng_ext(or_rng *r);
uint32_t or_rng_bits(or_rng *r, unsigned bits);

/* ---- Hal ops on host memory (hal.c); matrices are column-major [count][n] ---- */
void or_batch_interpolate_ntt(fp *io, size_t count, size_t n);
void or_batch_evaluate_ntt(fp *io, size_t count, size_t n, unsigned expand_bits);
void or_batch_expand_into_evaluate_ntt(fp *out, const fp *in, size_t count, size_t in_n, unsigned expand_bits);
void or_batch_bit_reverse(fp *io, size_t count, size_t n);
void or_zk_shift(fp *io, size_t count, size_t n);
void or_hash_rows(uint32_t *out_digests, const fp *matrix, size_t rows, size_t cols);
void or_hash_fold(uint32_t *nodes, size_t in_size, size_t out_size);
/* one row / one pair at a time through the scalar sponge: the definitions the lane-blocked forms above are tested against */
void or_hash_rows_scalar(uint32_t *out_digests, const fp *matrix, size_t rows, size_t cols);
void or_hash_fold_scalar(uint32_t *nodes, size_t in_size, size_t out_size);
void or_fri_fold(fp *out, const fp *in, const fp4 *mix, size_t out_count);
void or_batch_evaluate_any(const fp *coeffs, size_t poly_count, size_t n, const uint32_t *which, const fp4 *xs,
                           size_t eval_count, fp4 *out);
void or_mix_poly_coeffs(fp4 *out, const fp4 *mix_start, const fp4 *mix, const fp *in, const uint32_t *combos,
                        size_t input_size, size_t count);
void or_eltwise_add_elem(fp *out, const fp *a, const fp *b, size_t n);
void or_eltwise_sum_extelem(fp *out, const fp4 *in, size_t count, size_t to_add);
void or_eltwise_copy_elem(fp *out, const fp *in, size_t n);
void or_eltwise_zeroize_elem(fp *io, size_t n);
void or_gather_sample(fp *dst, const fp *src, size_t idx, size_t size, size_t stride);
void or_scatter(fp *into, const uint32_t *index, const uint32_t *offsets, const fp *values, size_t count);
void or_prefix_products(fp4 *io, size_t n);
/* vectorised field helpers for tests */
void or_fp_mul_vec(fp *out, const fp *a, const fp *b, size_t n);
void or_fp4_mul_vec(fp4 *out, const fp4 *a, const fp4 *b, size_t n);
void or_fp4_inv_vec(fp4 *out, const fp4 *a, size_t n);
void or_to_montgomery(uint32_t *io, size_t n);
void or_from_montgomery(uint32_t *io, size_t n);

/* ---- circuit (circuit.c): taps + constraint program loaded from a blob (format: DESIGN.md "circuit blob") ---- */
typedef struct or_circuit or_circuit;
or_circuit *or_circuit_load(const uint32_t *blob, size_t words);
void or_circuit_free(or_circuit *c);
size_t or_circuit_group_size(const or_circuit *c, int group);
size_t or_circuit_globals_size(const or_circuit *c);
size_t or_circuit_mix_size(const or_circuit *c);
/* CircuitHal::eval_check: groups = evaluated LDE matrices [cols][4N] in group order accum, code, data. */
void or_eval_check(const or_circuit *c, fp *check, const fp *const groups[3], const fp *globals, const fp *mix,
                   const fp4 *poly_mix, unsigned po2);
/* the same, one point at a time with untyped (fp4) values: the definition the lane-blocked or_eval_check is tested against */
void or_eval_check_scalar(const or_circuit *c, fp *check, const fp *const groups[3], const fp *globals, const fp *mix,
                          const fp4 *poly_mix, unsigned po2);
/* verifier-side poly_ext on ext tap evaluations */
fp4 or_poly_ext(const or_circuit *c, const fp4 *poly_mix, const fp4 *eval_u, const fp *globals, const fp *mix);
/* synthetic-circuit witness generation / accumulate (the "next row" f1 producer) */
void or_witgen(const or_circuit *c, unsigned po2, fp *code, fp *data, fp *globals, uint64_t seed, size_t zk_rows);
void or_accumulate(const or_circuit *c, unsigned po2, fp *accum, const fp *code, const fp *data, const fp *mix,
                   const fp *accum_noise, size_t zk_rows);

/* control roots: the verifier rejects (code 14) a seal whose code-group root differs from the registered one and
 * (code 15) a seal of a po2 that has none registered -- risc0-zkp verify/mod.rs check_code / the control-ID table */
void or_circuit_set_control_root(or_circuit *c, unsigned po2, const uint32_t *root8);
void or_control_root(unsigned po2, const fp *code, size_t cols, uint32_t *root8);

/* ---- prover / verifier (zkp.c) ---- */
/* Returns seal length in words (0 on failure); *seal is malloc'd. timings (optional, 16 doubles) gets per-phase seconds. */
size_t or_prove_segment(const or_circuit *c, unsigned po2, const fp *code, const fp *data, const fp *globals,
                        const fp *accum_noise, size_t zk_rows, uint32_t **seal, double *timings);
/* 0 = accepted; nonzero = rejection reason code */
int or_verify_segment(const or_circuit *c, const uint32_t *seal, size_t seal_words);
void or_free(void *p);
int or_num_threads(void);
void or_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
