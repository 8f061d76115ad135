/* protocol_params.h -- every protocol convention of the STARK segment proof that this repository had to RECALL.
 *
 * ONE copy, included by the oracle (oracle/oracle.h), the product prover (zeth_b200/csrc/prover.cpp, transcript.cpp,
 * circuit.cpp) and the product verifier (zeth_b200/csrc/verifier.cpp).  The arithmetic these describe lives in
 * un-vendored crates (risc0-zkp 3.0.2, /root/reference Cargo.lock:5393-5396) that cannot be read or built offline, so
 * each entry carries the confidence tag of SURVEY.md Appendix A.  A single real seal from risc0's own prover is enough
 * to flip any [RECALL med] line below: change it HERE and oracle, prover and verifier move together (the frozen
 * digests under tests/golden/ will then have to be regenerated with tools/make_golden.py).
 *
 * Plain C (C11 / C++17), no dependencies: macros and `static inline` helpers only.
 */
#ifndef B2Z_PROTOCOL_PARAMS_H
#define B2Z_PROTOCOL_PARAMS_H
#include <stddef.h>
#include <stdint.h>

/* ---- risc0-zkp src/lib.rs constants ---- */
#define PP_INV_RATE 4         /* [RECALL hi]  LDE blow-up */
#define PP_QUERIES 50         /* [RECALL hi]  FRI / Merkle queries per seal */
#define PP_FRI_FOLD 16        /* [RECALL hi]  fold factor per FRI round */
#define PP_FRI_MIN_DEGREE 256 /* [RECALL hi]  FRI stops when the polynomial has <= this many coefficients */
#define PP_EXT_SIZE 4         /* [RECALL hi]  F_p[x]/(x^4 - 11) */
#define PP_CHECK_SIZE 16      /* [RECALL hi]  INV_RATE * EXT_SIZE columns of the check group */
#define PP_DIGEST_WORDS 8     /* [RECALL hi] */
#define PP_MIN_CYCLES_PO2 13  /* [RECALL hi] */
#define PP_MAX_CYCLES_PO2 24  /* [RECALL hi] */
#define PP_ZK_CYCLES 1994     /* [RECALL med] rows of random noise at the end of every committed group */

/* [RECALL med] 16 bytes hashed (one field element per byte) into the transcript before anything else */
#define PP_PROOF_SYSTEM_INFO "RISC0_STARK:v1__"
#define PP_INFO_BYTES 16

/* ---- Poseidon2 sponge / rng framing (core/hash/poseidon2/{mod,rng}.rs) ---- */
#define PP_SPONGE_RATE 16     /* [RECALL med] overwrite-mode absorb of 16 cells, zero-padded tail, empty input = one permutation */
#define PP_SPONGE_OUT 8       /* [RECALL med] digest = cells 0..7 as raw Montgomery words */
/* [RECALL lo-med] Poseidon2Rng::random_bits: one draw, then ALWAYS this many more, each replacing a zero value */
#define PP_RANDOM_BITS_EXTRA_DRAWS 3

/* ---- Merkle tree (core/merkle.rs MerkleTreeParams) ---- */
/* [RECALL med] top_layer = largest l in [1, layers) with 2^l <= queries (0 when there is none); top_size = 2^top_layer.
 * With QUERIES = 50 this gives top_size = 32.  If upstream compares the other way round it is 64 -- a real seal's
 * length decides (seal words differ by 8 * 32 per committed tree). */
static inline size_t pp_merkle_top_layer(size_t layers, size_t queries) {
  size_t top_layer = 0;
  for (size_t i = 1; i < layers; i++) {
    if (((size_t)1 << i) > queries) break;
    top_layer = i;
  }
  return top_layer;
}

/* ---- transcript / seal order (prove/prover.rs, rv32im prove/mod.rs) ---- */
/* register-group numbering of the circuit: taps.rs REGISTER_GROUP_{ACCUM,CODE,DATA} [RECALL hi] */
#define PP_GROUP_ACCUM 0
#define PP_GROUP_CODE 1
#define PP_GROUP_DATA 2
/* [RECALL med] commit order: code, data, (mix draws), accum, (poly_mix), check.  Query openings per position: the three
 * register groups in numbering order accum, code, data, then check, then one row per FRI round. */
static const int PP_COMMIT_ORDER[3] = {PP_GROUP_CODE, PP_GROUP_DATA, PP_GROUP_ACCUM};
static const int PP_QUERY_ORDER[3] = {PP_GROUP_ACCUM, PP_GROUP_CODE, PP_GROUP_DATA};
/* [RECALL med] the check polynomial's 4N bit-reversed coefficients are read as quarters g_k, h(x) = sum g_k(x^4) x^k,
 * in quarter order k = 0, 2, 1, 3 (verify/mod.rs `remap`) */
static const int PP_CHECK_REMAP[4] = {0, 2, 1, 3};
/* [RECALL med] the header written first into the seal: globals, then po2 as one field element */
#define PP_HEADER_EXTRA_WORDS 1

/* ---- synthetic sessions: OUR convention, not upstream's (the rv32im globals layout is a generated artifact that is
 * unavailable offline).  A circuit with >= 24 globals carries a receipt claim in its last 16 globals: the pre-state
 * digest in globals[G-16, G-8) and the post-state digest in globals[G-8, G).  The synthetic witness generators derive
 * post(seed) == pre(seed + 1), so a session of consecutive seeds forms a valid claim chain (receipt.verify walks it). */
#define PP_CLAIM_DIGEST_WORDS 8
static inline long pp_claim_pre_offset(size_t globals_size) { return globals_size >= 24 ? (long)globals_size - 16 : -1; }
static inline long pp_claim_post_offset(size_t globals_size) { return globals_size >= 24 ? (long)globals_size - 8 : -1; }

#endif
