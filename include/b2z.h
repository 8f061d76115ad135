/* b2z.h -- C ABI of the B200-native STARK segment prover (libb2z.so).
 *
 * Drop-in boundary (seam S3): zeth reaches the prover through ONE call,
 *     default_prover().prove(env, elf)            /root/reference/crates/host/src/lib.rs:137
 * (configured at lib.rs:132-136, result checked at crates/host/src/bin/cli.rs:103-107).  Below that call the
 * un-vendored risc0 crates drive a `Hal` / `CircuitHal` back-end whose CUDA implementation crosses an
 * `extern "C"` boundary into risc0-sys kernels.  This header is that boundary for the B200 back-end: plain
 * pointers and sizes, no C++ / torch types.  Each entry point names the upstream trait method it replaces
 * (risc0-zkp 3.0.2 `hal/mod.rs` trait Hal, the `prove/` modules; risc0-circuit-rv32im 4.0.2 CircuitHal), which is what a
 * Rust `impl Hal for B200Hal` shim binds.
 *
 * Conventions (same as upstream's FFI): every function returns NULL on success or a NUL-terminated error string.
 * OWNERSHIP DIFFERENCE from upstream: risc0-sys returns a heap string that the Rust side copies and frees; here the
 * string is owned by the library (thread-local, valid until the same thread's next failing call) -- the caller copies it
 * and must NOT free it (INTEGRATION.md shows the Rust `CStr::from_ptr(..).to_owned()` side).  Field elements are u32 Montgomery words < P = 15*2^27+1; matrices
 * are column-major [cols][rows]; ext elements are 4 consecutive words; digests are 8 words.  All `dev` pointers
 * are device pointers of the context's device; work is enqueued on the context's stream and is asynchronous
 * unless stated.  A context is single-threaded (upstream `Rc<dyn Prover>` is !Send): one context per thread; it may be
 * created on one thread and used on another (zeth builds the prover inside spawn_blocking, lib.rs:131): every entry
 * point re-asserts the context's device on the calling thread.
 * There is NO CPU fallback: without a usable CUDA device every compute entry point returns an error.
 */
#ifndef B2Z_H
#define B2Z_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct b2z_ctx b2z_ctx;
typedef struct b2z_circuit b2z_circuit;
typedef struct b2z_session b2z_session;
typedef const char* b2z_err;

/* protocol constants (risc0-zkp 3.0.2 src/lib.rs): single source of truth with confidence tags in protocol_params.h */
#include "protocol_params.h"
#define B2Z_INV_RATE PP_INV_RATE
#define B2Z_QUERIES PP_QUERIES
#define B2Z_FRI_FOLD PP_FRI_FOLD
#define B2Z_FRI_MIN_DEGREE PP_FRI_MIN_DEGREE
#define B2Z_EXT_SIZE PP_EXT_SIZE
#define B2Z_CHECK_SIZE PP_CHECK_SIZE
#define B2Z_DIGEST_WORDS PP_DIGEST_WORDS
#define B2Z_ZK_CYCLES PP_ZK_CYCLES
#define B2Z_GROUP_ACCUM PP_GROUP_ACCUM
#define B2Z_GROUP_CODE PP_GROUP_CODE
#define B2Z_GROUP_DATA PP_GROUP_DATA

/* ---- context / memory: replaces CudaHal::new + Buffer alloc / copy_from / view (hal/cuda.rs) ---- */
const char* b2z_version(void);
b2z_err b2z_init(int device, b2z_ctx** out);
b2z_err b2z_destroy(b2z_ctx* ctx);
/* adopt an externally owned cudaStream_t (e.g. torch.cuda.Stream().cuda_stream); NULL restores the own stream */
b2z_err b2z_set_stream(b2z_ctx* ctx, void* cuda_stream);
b2z_err b2z_sync(b2z_ctx* ctx);
b2z_err b2z_alloc(b2z_ctx* ctx, size_t bytes, void** dev);
b2z_err b2z_free(b2z_ctx* ctx, void* dev);
b2z_err b2z_h2d(b2z_ctx* ctx, void* dev, const void* host, size_t bytes);
b2z_err b2z_d2h(b2z_ctx* ctx, void* host, const void* dev, size_t bytes); /* synchronises */
/* number of kernels this context has launched since creation (bench.py `gpu_launches`) */
uint64_t b2z_launch_count(const b2z_ctx* ctx);

/* ---- Hal operations (trait Hal, risc0-zkp hal/mod.rs; CPU semantics in hal/cpu.rs) ---- */
/* Hal::batch_interpolate_ntt(io, count): natural evaluations -> bit-reversed coefficients, scaled by 1/n */
b2z_err b2z_batch_interpolate_ntt(b2z_ctx* ctx, uint32_t* io, size_t count, unsigned po2);
/* Hal::batch_evaluate_ntt(io, count, expand_bits): bit-reversed coefficients -> natural evaluations */
b2z_err b2z_batch_evaluate_ntt(b2z_ctx* ctx, uint32_t* io, size_t count, unsigned po2, unsigned expand_bits);
/* Hal::batch_expand_into_evaluate_ntt(out, in, count, expand_bits): out[i] = in[i >> bits], remaining layers */
b2z_err b2z_batch_expand_into_evaluate_ntt(b2z_ctx* ctx, uint32_t* out, const uint32_t* in, size_t count, unsigned in_po2,
                                           unsigned expand_bits);
/* Hal::batch_bit_reverse(io, count) */
b2z_err b2z_batch_bit_reverse(b2z_ctx* ctx, uint32_t* io, size_t count, unsigned po2);
/* Hal::zk_shift(io, count): bit-reversed slot j *= 3^rev(j) */
b2z_err b2z_zk_shift(b2z_ctx* ctx, uint32_t* io, size_t count, unsigned po2);
/* Hal::hash_rows(output, matrix): Poseidon2 sponge of every row of a column-major [cols][rows] matrix */
b2z_err b2z_hash_rows(b2z_ctx* ctx, uint32_t* out_digests, const uint32_t* matrix, size_t rows, size_t cols);
/* The same sponge absorbed in column chunks (b2z_prove_segment hashes each witness chunk as it lands over PCIe):
 * call with the chunks of the matrix in order, carry_in = 0 on the first, last = 1 on the final one; every non-final
 * chunk must be a positive multiple of 16 columns.  `out_digests` (8 words per row) holds the carried sponge capacity
 * between calls and the digests after the last call -- identical to one b2z_hash_rows over all the columns. */
b2z_err b2z_hash_rows_chunk(b2z_ctx* ctx, uint32_t* out_digests, const uint32_t* matrix_chunk, size_t rows, size_t cols, int carry_in,
                            int last);
/* Hal::hash_fold(io, input_size, output_size): node[out+j] = hash_pair(node[in+2j], node[in+2j+1]) */
b2z_err b2z_hash_fold(b2z_ctx* ctx, uint32_t* nodes, size_t in_size, size_t out_size);
/* MerkleTreeProver::new (prove/merkle.rs): hash_rows into nodes[rows..2rows) then every hash_fold layer */
b2z_err b2z_merkle_build(b2z_ctx* ctx, uint32_t* nodes, const uint32_t* matrix, size_t rows, size_t cols);
/* Hal::fri_fold(output, input, mix): mix = 4 host words */
b2z_err b2z_fri_fold(b2z_ctx* ctx, uint32_t* out, const uint32_t* in, const uint32_t* mix_host, size_t out_count);
/* Hal::batch_evaluate_any(coeffs, poly_count, which, xs, out): which/xs/out are device arrays */
b2z_err b2z_batch_evaluate_any(b2z_ctx* ctx, const uint32_t* coeffs, size_t poly_count, unsigned po2, const uint32_t* which,
                               const uint32_t* xs, size_t eval_count, uint32_t* out);
/* Hal::mix_poly_coeffs(output, mix_start, mix, input, combos, input_size, count); mix values are host words */
b2z_err b2z_mix_poly_coeffs(b2z_ctx* ctx, uint32_t* out, const uint32_t* mix_start_host, const uint32_t* mix_host,
                            const uint32_t* in, const uint32_t* combos_dev, size_t input_size, size_t count);
/* Hal::eltwise_add_elem / eltwise_sum_extelem / eltwise_copy_elem / eltwise_zeroize_elem */
b2z_err b2z_eltwise_add_elem(b2z_ctx* ctx, uint32_t* out, const uint32_t* a, const uint32_t* b, size_t n);
b2z_err b2z_eltwise_sum_extelem(b2z_ctx* ctx, uint32_t* out, const uint32_t* in, size_t count, size_t to_add);
b2z_err b2z_eltwise_copy_elem(b2z_ctx* ctx, uint32_t* out, const uint32_t* in, size_t n);
b2z_err b2z_eltwise_zeroize_elem(b2z_ctx* ctx, uint32_t* io, size_t n);
/* Hal::gather_sample(dst, src, idx, size, stride) */
b2z_err b2z_gather_sample(b2z_ctx* ctx, uint32_t* dst, const uint32_t* src, size_t idx, size_t size, size_t stride);
/* Hal::scatter(into, index, offsets, values): CSR form -- `index` holds count + 1 starts; cycle c performs
 * into[offsets[i]] = values[i] for i in [index[c], index[c+1]) in order.  All four are device arrays. */
b2z_err b2z_scatter(b2z_ctx* ctx, uint32_t* into, const uint32_t* index, const uint32_t* offsets, const uint32_t* values, size_t count);
/* Hal::prefix_products(io): in-place inclusive prefix product of `n` AoS ext elements (4 words each) */
b2z_err b2z_prefix_products(b2z_ctx* ctx, uint32_t* ext_io, size_t n);
/* core/poly.rs poly_divide applied to an ext coefficient array on the device (prover.rs runs it on the CPU):
 * in-place synthetic division of `count` AoS ext coefficients by (X - z); remainder -> 4 host words */
b2z_err b2z_poly_divide(b2z_ctx* ctx, uint32_t* ext_coeffs, size_t count, const uint32_t* z_host, uint32_t* remainder_host);

/* ---- circuit: taps + constraint program (risc0-zkp taps.rs / adapter.rs PolyExtStep), as a blob ---- */
b2z_err b2z_circuit_load(b2z_ctx* ctx, const uint32_t* blob, size_t words, b2z_circuit** out);
/* CPU-side build step: generate + NVRTC-compile the circuit kernels into the on-disk cubin cache (no GPU needed) */
b2z_err b2z_circuit_precompile(const uint32_t* blob, size_t words);
b2z_err b2z_circuit_free(b2z_circuit* c);
size_t b2z_circuit_group_size(const b2z_circuit* c, int group);
size_t b2z_circuit_globals_size(const b2z_circuit* c);
size_t b2z_circuit_mix_size(const b2z_circuit* c);
size_t b2z_circuit_seal_words(const b2z_circuit* c, unsigned po2);
/* generated CUDA source of the circuit kernels (NUL-terminated; owned by the circuit) */
const char* b2z_circuit_source(const b2z_circuit* c);
/* CircuitHal::eval_check(check, groups, globals, poly_mix, po2, steps): groups = device LDE matrices in order
 * accum, code, data; globals/mix device arrays; poly_mix = 4 host words */
b2z_err b2z_eval_check(b2z_ctx* ctx, const b2z_circuit* c, uint32_t* check, const uint32_t* const groups_dev[3],
                       const uint32_t* globals_dev, const uint32_t* mix_dev, const uint32_t* poly_mix_host, unsigned po2);
/* CircuitHal::accumulate (synthetic-circuit form: ext running sums fed by the Fiat-Shamir mix) */
b2z_err b2z_accumulate(b2z_ctx* ctx, const b2z_circuit* c, unsigned po2, uint32_t* accum, const uint32_t* code,
                       const uint32_t* data, const uint32_t* mix_dev, const uint32_t* accum_noise_dev, size_t zk_rows);
/* witness generation for synthetic circuits (the reference's witgen / step_exec stage): fills code, data, globals */
b2z_err b2z_witgen(b2z_ctx* ctx, const b2z_circuit* c, unsigned po2, uint32_t* code, uint32_t* data, uint32_t* globals_dev,
                   uint32_t* accum_noise_dev, uint64_t seed, size_t zk_rows);

/* ---- segment prover: replaces SegmentProver::prove -> seal: Vec<u32> (risc0-circuit-rv32im prove/) ---- */
/* Witness already resident in HBM (device pointers). seal_host must hold b2z_circuit_seal_words() words. Synchronises. */
b2z_err b2z_prove_segment_device(b2z_ctx* ctx, const b2z_circuit* c, unsigned po2, const uint32_t* code_dev,
                                 const uint32_t* data_dev, const uint32_t* globals_host, const uint32_t* accum_noise_dev,
                                 size_t zk_rows, uint32_t* seal_host, size_t* seal_words);
/* Host buffers in, seal out: H2D of the witness and D2H of the seal happen inside the call, the upload overlapped with
 * the commits.  Pinned buffers (cudaMallocHost / cudaHostRegister) are copied directly; PAGEABLE buffers (a Rust
 * Vec<u32>, a numpy array) are detected and staged through a pinned ring by helper threads, so both overlap.
 * A witness that does not satisfy the circuit is an ERROR, not a seal (check-polynomial degree test + exact DEEP
 * divisions); so is it for b2z_prove_segment_device. */
b2z_err b2z_prove_segment(b2z_ctx* ctx, const b2z_circuit* c, unsigned po2, const uint32_t* code_host, const uint32_t* data_host,
                          const uint32_t* globals_host, const uint32_t* accum_noise_host, size_t zk_rows, uint32_t* seal_host,
                          size_t* seal_words);
/* Control roots (risc0: the control-ID table): the Merkle root that the commitment of a program's code group has at a
 * given po2.  verify/mod.rs check_code(po2, code_root) binds a seal to the program -- the check behind
 * receipt.verify(image_id), /root/reference crates/host/src/bin/cli.rs:103.  b2z_compute_control_root derives it on the
 * GPU from the code columns; b2z_circuit_set_control_root registers it with a circuit (verifier side). */
b2z_err b2z_compute_control_root(b2z_ctx* ctx, unsigned po2, const uint32_t* code_dev, size_t code_cols, uint32_t* root8_host);
b2z_err b2z_circuit_set_control_root(b2z_circuit* c, unsigned po2, const uint32_t* root8);
int b2z_circuit_get_control_root(const b2z_circuit* c, unsigned po2, uint32_t* root8); /* 1 if registered */

/* Per-segment STARK verifier on the host CPU: what `receipt.verify(image_id)` (/root/reference crates/host/src/bin/cli.rs:103)
 * runs for each SegmentReceipt seal (risc0-zkp verify/mod.rs).  Needs no GPU (load the circuit with ctx == NULL).
 * Returns 0 if the seal is accepted, a positive rejection code otherwise, -1 on bad arguments.  Codes: 1-13 transcript /
 * Merkle / FRI failures, 14 = the seal's code root differs from the registered control root, 15 = no control root is
 * registered for the seal's po2 (fails closed). */
int b2z_verify_segment(const b2z_circuit* c, const uint32_t* seal, size_t seal_words);

/* ---- session level: Prover::prove -> Receipt, Receipt::verify (risc0-zkvm 3.0.3; /root/reference
 * crates/host/src/lib.rs:131-142 and crates/host/src/bin/cli.rs:103-107) ---- */
/* One segment of a session = the cached witness of SegmentProver::prove's input (noise is an input). */
typedef struct b2z_segment_input {
  uint32_t index;              /* position in the session (receipts come back in this order) */
  uint32_t po2;
  const uint32_t* code;        /* [code cols][2^po2] */
  const uint32_t* data;        /* [data cols][2^po2] */
  const uint32_t* globals;     /* host, always */
  const uint32_t* accum_noise; /* [accum cols][zk_rows] */
  size_t zk_rows;
  int on_device;               /* 0: code/data/accum_noise are host pointers (pinned or pageable); 1: device pointers */
} b2z_segment_input;
/* SegmentReceipt{seal, index, hashfn, verifier_parameters, claim} */
typedef struct b2z_segment_receipt {
  uint32_t index;
  uint32_t po2;
  char hashfn[16];                 /* "poseidon2" */
  uint32_t verifier_parameters[8]; /* digest naming (proof system, circuit) */
  uint32_t claim_pre[8];           /* ReceiptClaim stand-in: pre / post state digests (protocol_params.h, synthetic sessions) */
  uint32_t claim_post[8];
  uint32_t* seal;                  /* owned by the receipt; the bit-exact Vec<u32> transcript */
  size_t seal_words;
} b2z_segment_receipt;
/* Receipt{inner: Composite{segments, ..}, journal} */
typedef struct b2z_receipt {
  b2z_segment_receipt* segments;
  size_t n_segments;
  uint32_t verifier_parameters[8];
  uint8_t* journal;
  size_t journal_len;
} b2z_receipt;
/* A prover server for one circuit on one GPU with `in_flight` (1..8) segments proved concurrently on separate streams. */
b2z_err b2z_session_create(int device, const uint32_t* blob, size_t words, unsigned in_flight, b2z_session** out);
b2z_err b2z_session_destroy(b2z_session* s);
b2z_circuit* b2z_session_circuit(b2z_session* s);               /* worker 0's circuit (seal sizes, verification) */
b2z_ctx* b2z_session_context(b2z_session* s, unsigned worker);  /* a worker's context (witness generation, buffers) */
unsigned b2z_session_in_flight(const b2z_session* s);
b2z_err b2z_session_set_control_root(b2z_session* s, unsigned po2, const uint32_t* root8);
/* ProverServer::prove_session: every segment proved (one retry on a device error; an unsatisfied witness fails the
 * session), SegmentReceipts assembled in index order.  Free the result with b2z_receipt_free. */
b2z_err b2z_prove_session(b2z_session* s, const b2z_segment_input* segments, size_t n_segments, const uint8_t* journal,
                          size_t journal_len, b2z_receipt** out);
/* Seal cache / resume: with a directory set, proved seals are stored there keyed by (circuit, po2, index, header digest) and
 * a later session loads them instead of proving again (re-verified first when a control root is registered).  NULL / ""
 * switches it off.  The reference's analogue is its input cache, /root/reference crates/host/src/bin/cli.rs:113-145. */
b2z_err b2z_session_set_seal_cache(b2z_session* s, const char* dir);
size_t b2z_session_last_cache_hits(const b2z_session* s);
uint64_t b2z_session_launch_count(const b2z_session* s);
double b2z_session_last_wall_ms(const b2z_session* s);
size_t b2z_session_last_retries(const b2z_session* s);
size_t b2z_session_last_segment_ms(const b2z_session* s, double* ms, size_t cap);
/* Receipt containers: assemble from seals (claims are decoded from each seal's header), free, (de)serialise. */
b2z_err b2z_receipt_assemble(const b2z_circuit* c, const uint32_t* const* seals, const size_t* seal_words, size_t n_segments,
                             const uint8_t* journal, size_t journal_len, b2z_receipt** out);
b2z_err b2z_receipt_free(b2z_receipt* r);
b2z_err b2z_receipt_serialize(const b2z_receipt* r, uint8_t** bytes, size_t* len); /* free with b2z_bytes_free */
b2z_err b2z_receipt_deserialize(const uint8_t* bytes, size_t len, b2z_receipt** out);
void b2z_bytes_free(uint8_t* bytes);
/* Receipt::verify(image_id): segments in index order, every seal verified (incl. control root), each claim equal to what
 * its seal's header says, claims chained (post of i == pre of i + 1), first pre-state == image_id (NULL skips that).
 * 0 = accepted; 1-15 = b2z_verify_segment code of segment *bad_segment; 101 empty, 102 index order, 103 hashfn,
 * 104 verifier parameters, 105 claim != seal, 106 claim chain broken, 107 image id mismatch. */
int b2z_verify_receipt(const b2z_circuit* c, const b2z_receipt* r, const uint32_t* image_id8, size_t* bad_segment);
/* per-phase device milliseconds of the last prove call on this context (names via b2z_phase_name; returns count) */
size_t b2z_last_phase_ms(const b2z_ctx* ctx, double* ms, size_t cap);
const char* b2z_phase_name(size_t i);

/* host-side Poseidon2 (transcript) helpers, exported for tests: hash_elem_slice / hash_pair / permutation */
void b2z_host_poseidon2_mix(uint32_t* state24);
/* the same permutation computed by the exact-FP64 instruction stream the mixed hash kernels co-schedule (host copy) */
void b2z_host_poseidon2_mix_f64(uint32_t* state24);
void b2z_host_hash_elem_slice(const uint32_t* elems, size_t n, uint32_t* digest8);

#ifdef __cplusplus
}
#endif
#endif
