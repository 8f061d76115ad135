// extern "C" surface declared in include/b2z.h.
#include <cstring>

#include "circuit.h"
#include "common.h"

namespace b2z {
b2z_err merkle_build(b2z_ctx* ctx, uint32_t* nodes, const uint32_t* matrix, size_t rows, size_t cols);
struct WitnessReady;
b2z_err prove_segment_device(b2z_ctx* ctx, const b2z_circuit* c, unsigned po2, const uint32_t* code_dev, const uint32_t* data_dev,
                             const uint32_t* globals_host, const uint32_t* accum_noise_dev, size_t zk_rows, uint32_t* seal_host,
                             size_t* seal_words_out, const WitnessReady* ready);
b2z_err prove_segment_host(b2z_ctx* ctx, const b2z_circuit* c, unsigned po2, const uint32_t* code_host, const uint32_t* data_host,
                           const uint32_t* globals_host, const uint32_t* accum_noise_host, size_t zk_rows, uint32_t* seal_host,
                           size_t* seal_words);
b2z_err eval_check_launch(b2z_ctx* ctx, const b2z_circuit* c, uint32_t* check, const uint32_t* const groups[3], const uint32_t* globals_host,
                          const uint32_t* mix_host, const Fp4& poly_mix, unsigned po2);
b2z_err witgen_launch(b2z_ctx* ctx, const b2z_circuit* c, unsigned po2, uint32_t* code, uint32_t* data, uint32_t* globals_dev,
                      uint32_t* accum_noise_dev, uint64_t seed, size_t zk_rows);
const char* phase_name(size_t i);
void ntt_release_tables(b2z_ctx* ctx);
int verify_segment(const b2z_circuit* c, const uint32_t* seal, size_t seal_words);
b2z_err compute_control_root(b2z_ctx* ctx, unsigned po2, const uint32_t* code_dev, size_t cols, uint32_t* root_host);
}  // namespace b2z
using namespace b2z;

static Fp4 ext_from(const uint32_t* w) { return fp4_make(w[0], w[1], w[2], w[3]); }

extern "C" {

const char* b2z_version(void) { return "b2z 0.1 (sm_100a)"; }

b2z_err b2z_init(int device, b2z_ctx** out) {
  B2Z_REQUIRE(out != nullptr, "null out pointer");
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return set_error(std::string("no usable CUDA device (this library has no CPU fallback): ") + cudaGetErrorString(e));
  B2Z_REQUIRE(device >= 0 && device < count, "device index out of range");
  B2Z_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  B2Z_CUDA(cudaGetDeviceProperties(&prop, device));
  B2Z_REQUIRE(prop.major >= 10, "sm_100a kernels need a Blackwell-class GPU");
  b2z_ctx* ctx = new b2z_ctx();
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  B2Z_CUDA(cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking));
  ctx->stream = ctx->own_stream;
  cudaMemPool_t pool;
  B2Z_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
  uint64_t thresh = UINT64_MAX;
  B2Z_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thresh));
  *out = ctx;
  return nullptr;
}
b2z_err b2z_destroy(b2z_ctx* ctx) {
  if (!ctx) return nullptr;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  ntt_release_tables(ctx);
  for (int i = 0; i <= MAX_PO2; i++)
    if (ctx->pow3rev[i]) cudaFree(ctx->pow3rev[i]);
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  if (ctx->stage_ring) cudaFreeHost(ctx->stage_ring);
  for (auto e : ctx->stage_events) cudaEventDestroy(e);
  if (ctx->hash_sched) cudaFree(ctx->hash_sched);
  for (auto& pr : ctx->phase_events) {
    cudaEventDestroy(pr.first);
    cudaEventDestroy(pr.second);
  }
  for (auto e : ctx->h2d_events) cudaEventDestroy(e);
  for (auto e : ctx->event_pool) cudaEventDestroy(e);
  if (ctx->wait_event) cudaEventDestroy(ctx->wait_event);
  if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
  if (ctx->aux_stream) cudaStreamDestroy(ctx->aux_stream);
  if (ctx->aux_fork) cudaEventDestroy(ctx->aux_fork);
  if (ctx->aux_join) cudaEventDestroy(ctx->aux_join);
  cudaStreamDestroy(ctx->own_stream);
  delete ctx;
  return nullptr;
}
b2z_err b2z_set_stream(b2z_ctx* ctx, void* s) {
  B2Z_REQUIRE(ctx, "null context");
  B2Z_CUDA(cudaSetDevice(ctx->device));
  B2Z_CUDA(cudaStreamSynchronize(ctx->stream));
  ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
  return nullptr;
}
b2z_err b2z_sync(b2z_ctx* ctx) {
  B2Z_REQUIRE(ctx, "null context");
  B2Z_CUDA(cudaSetDevice(ctx->device));
  B2Z_CUDA(cudaStreamSynchronize(ctx->stream));
  return nullptr;
}
b2z_err b2z_alloc(b2z_ctx* ctx, size_t bytes, void** dev) {
  B2Z_REQUIRE(ctx && dev, "null argument");
  B2Z_CUDA(cudaSetDevice(ctx->device));
  B2Z_CUDA(cudaMallocAsync(dev, bytes ? bytes : 4, ctx->stream));
  return nullptr;
}
b2z_err b2z_free(b2z_ctx* ctx, void* dev) {
  B2Z_REQUIRE(ctx, "null context");
  B2Z_CUDA(cudaSetDevice(ctx->device));
  if (dev) B2Z_CUDA(cudaFreeAsync(dev, ctx->stream));
  return nullptr;
}
b2z_err b2z_h2d(b2z_ctx* ctx, void* dev, const void* host, size_t bytes) {
  B2Z_REQUIRE(ctx, "null context");
  B2Z_CUDA(cudaSetDevice(ctx->device));
  B2Z_CUDA(cudaMemcpyAsync(dev, host, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return nullptr;
}
b2z_err b2z_d2h(b2z_ctx* ctx, void* host, const void* dev, size_t bytes) {
  B2Z_REQUIRE(ctx, "null context");
  B2Z_CUDA(cudaSetDevice(ctx->device));
  B2Z_CUDA(cudaMemcpyAsync(host, dev, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  B2Z_CUDA(cudaStreamSynchronize(ctx->stream));
  return nullptr;
}
uint64_t b2z_launch_count(const b2z_ctx* ctx) { return ctx ? ctx->launches : 0; }

// every entry point re-asserts the context's device: the calling thread may not be the one that ran b2z_init
#define CTX_GUARD()                                                                          \
  do {                                                                                       \
    B2Z_REQUIRE(ctx != nullptr, "null context (b2z_init failed? there is no CPU fallback)"); \
    B2Z_CUDA(cudaSetDevice(ctx->device));                                                    \
  } while (0)

b2z_err b2z_batch_interpolate_ntt(b2z_ctx* ctx, uint32_t* io, size_t count, unsigned po2) {
  CTX_GUARD();
  return launch_interpolate(ctx, io, count, po2, false);
}
b2z_err b2z_batch_evaluate_ntt(b2z_ctx* ctx, uint32_t* io, size_t count, unsigned po2, unsigned expand_bits) {
  CTX_GUARD();
  return launch_evaluate(ctx, io, io, count, po2, expand_bits);
}
b2z_err b2z_batch_expand_into_evaluate_ntt(b2z_ctx* ctx, uint32_t* out, const uint32_t* in, size_t count, unsigned in_po2,
                                           unsigned expand_bits) {
  CTX_GUARD();
  B2Z_REQUIRE(out != in, "expand needs distinct buffers");
  return launch_evaluate(ctx, out, in, count, in_po2 + expand_bits, expand_bits);
}
b2z_err b2z_batch_bit_reverse(b2z_ctx* ctx, uint32_t* io, size_t count, unsigned po2) {
  CTX_GUARD();
  return launch_bit_reverse(ctx, io, count, po2);
}
b2z_err b2z_zk_shift(b2z_ctx* ctx, uint32_t* io, size_t count, unsigned po2) {
  CTX_GUARD();
  return launch_zk_shift(ctx, io, count, po2);
}
b2z_err b2z_hash_rows(b2z_ctx* ctx, uint32_t* out, const uint32_t* matrix, size_t rows, size_t cols) {
  CTX_GUARD();
  return launch_hash_rows(ctx, out, matrix, rows, cols);
}
b2z_err b2z_hash_rows_chunk(b2z_ctx* ctx, uint32_t* out, const uint32_t* matrix, size_t rows, size_t cols, int carry_in, int last) {
  CTX_GUARD();
  return launch_hash_rows_chunk(ctx, out, matrix, rows, cols, carry_in != 0, last != 0);
}
b2z_err b2z_hash_fold(b2z_ctx* ctx, uint32_t* nodes, size_t in_size, size_t out_size) {
  CTX_GUARD();
  return launch_hash_fold(ctx, nodes, in_size, out_size);
}
b2z_err b2z_merkle_build(b2z_ctx* ctx, uint32_t* nodes, const uint32_t* matrix, size_t rows, size_t cols) {
  CTX_GUARD();
  B2Z_REQUIRE(rows && (rows & (rows - 1)) == 0, "rows must be a power of two");
  return merkle_build(ctx, nodes, matrix, rows, cols);
}
b2z_err b2z_fri_fold(b2z_ctx* ctx, uint32_t* out, const uint32_t* in, const uint32_t* mix_host, size_t out_count) {
  CTX_GUARD();
  return launch_fri_fold(ctx, out, in, ext_from(mix_host), out_count);
}
b2z_err b2z_batch_evaluate_any(b2z_ctx* ctx, const uint32_t* coeffs, size_t poly_count, unsigned po2, const uint32_t* which,
                               const uint32_t* xs, size_t eval_count, uint32_t* out) {
  CTX_GUARD();
  (void)poly_count;
  return launch_evaluate_any(ctx, coeffs, po2, which, xs, eval_count, out);
}
b2z_err b2z_mix_poly_coeffs(b2z_ctx* ctx, uint32_t* out, const uint32_t* mix_start_host, const uint32_t* mix_host, const uint32_t* in,
                            const uint32_t* combos_dev, size_t input_size, size_t count) {
  CTX_GUARD();
  return launch_mix_poly_coeffs(ctx, out, ext_from(mix_start_host), ext_from(mix_host), in, combos_dev, input_size, count);
}
b2z_err b2z_eltwise_add_elem(b2z_ctx* ctx, uint32_t* out, const uint32_t* a, const uint32_t* b, size_t n) {
  CTX_GUARD();
  return launch_add_elem(ctx, out, a, b, n);
}
b2z_err b2z_eltwise_sum_extelem(b2z_ctx* ctx, uint32_t* out, const uint32_t* in, size_t count, size_t to_add) {
  CTX_GUARD();
  return launch_sum_extelem(ctx, out, in, count, to_add);
}
b2z_err b2z_eltwise_copy_elem(b2z_ctx* ctx, uint32_t* out, const uint32_t* in, size_t n) {
  CTX_GUARD();
  B2Z_CUDA(cudaMemcpyAsync(out, in, 4 * n, cudaMemcpyDeviceToDevice, ctx->stream));
  return nullptr;
}
b2z_err b2z_eltwise_zeroize_elem(b2z_ctx* ctx, uint32_t* io, size_t n) {
  CTX_GUARD();
  return launch_zeroize(ctx, io, n);
}
b2z_err b2z_gather_sample(b2z_ctx* ctx, uint32_t* dst, const uint32_t* src, size_t idx, size_t size, size_t stride) {
  CTX_GUARD();
  return launch_gather_sample(ctx, dst, src, idx, size, stride);
}
b2z_err b2z_scatter(b2z_ctx* ctx, uint32_t* into, const uint32_t* index, const uint32_t* offsets, const uint32_t* values, size_t count) {
  CTX_GUARD();
  return launch_scatter(ctx, into, index, offsets, values, count);
}
b2z_err b2z_prefix_products(b2z_ctx* ctx, uint32_t* ext_io, size_t n) {
  CTX_GUARD();
  return launch_prefix_products(ctx, ext_io, n);
}
b2z_err b2z_poly_divide(b2z_ctx* ctx, uint32_t* ext_coeffs, size_t count, const uint32_t* z_host, uint32_t* remainder_host) {
  CTX_GUARD();
  uint32_t* rem = nullptr;
  B2Z_CUDA(cudaMallocAsync(&rem, 16, ctx->stream));
  B2Z_TRY(launch_poly_divide(ctx, ext_coeffs, count, ext_from(z_host), rem));
  B2Z_CUDA(cudaMemcpyAsync(remainder_host, rem, 16, cudaMemcpyDeviceToHost, ctx->stream));
  B2Z_CUDA(cudaStreamSynchronize(ctx->stream));
  B2Z_CUDA(cudaFreeAsync(rem, ctx->stream));
  return nullptr;
}

b2z_err b2z_circuit_load(b2z_ctx* ctx, const uint32_t* blob, size_t words, b2z_circuit** out) {
  B2Z_REQUIRE(blob && out, "null argument");
  b2z_circuit* c = new b2z_circuit();
  const char* err = circuit_parse(c, blob, words);
  if (!err) {
    circuit_codegen(c);
    if (ctx) err = circuit_compile(ctx, c);  // ctx == NULL: parse + codegen only (CPU-side build step)
  }
  if (err) {
    delete c;
    return err;
  }
  *out = c;
  return nullptr;
}
// CPU-side: compile the circuit kernels into the on-disk cubin cache without touching a GPU (used by build())
b2z_err b2z_circuit_precompile(const uint32_t* blob, size_t words) {
  b2z_circuit c;
  B2Z_TRY(circuit_parse(&c, blob, words));
  circuit_codegen(&c);
  std::vector<char> cubin;
  return circuit_cubin(&c, cubin);
}
b2z_err b2z_circuit_free(b2z_circuit* c) {
  if (!c) return nullptr;
  if (c->library) cudaLibraryUnload(c->library);
  if (c->accum_col_dev) cudaFree(c->accum_col_dev);
  if (c->col_kind_dev) cudaFree(c->col_kind_dev);
  delete c;
  return nullptr;
}
size_t b2z_circuit_group_size(const b2z_circuit* c, int g) { return (c && g >= 0 && g < 3) ? c->group_size[g] : 0; }
size_t b2z_circuit_globals_size(const b2z_circuit* c) { return c ? c->globals_size : 0; }
size_t b2z_circuit_mix_size(const b2z_circuit* c) { return c ? c->mix_size : 0; }
size_t b2z_circuit_seal_words(const b2z_circuit* c, unsigned po2) { return c ? circuit_seal_words(c, po2) : 0; }
const char* b2z_circuit_source(const b2z_circuit* c) { return c ? c->source.c_str() : ""; }

b2z_err b2z_eval_check(b2z_ctx* ctx, const b2z_circuit* c, uint32_t* check, const uint32_t* const groups_dev[3],
                       const uint32_t* globals_dev, const uint32_t* mix_dev, const uint32_t* poly_mix_host, unsigned po2) {
  CTX_GUARD();
  B2Z_REQUIRE(c, "null circuit");
  // globals / mix arrive as device arrays (Hal buffers); the kernel takes them by value, so fetch the few words
  std::vector<uint32_t> g(c->globals_size + 1), m(c->mix_size + 1);
  if (c->globals_size) B2Z_CUDA(cudaMemcpyAsync(g.data(), globals_dev, 4 * c->globals_size, cudaMemcpyDeviceToHost, ctx->stream));
  if (c->mix_size) B2Z_CUDA(cudaMemcpyAsync(m.data(), mix_dev, 4 * c->mix_size, cudaMemcpyDeviceToHost, ctx->stream));
  B2Z_CUDA(cudaStreamSynchronize(ctx->stream));
  return eval_check_launch(ctx, c, check, groups_dev, g.data(), m.data(), ext_from(poly_mix_host), po2);
}
b2z_err b2z_accumulate(b2z_ctx* ctx, const b2z_circuit* c, unsigned po2, uint32_t* accum, const uint32_t* code, const uint32_t* data,
                       const uint32_t* mix_dev, const uint32_t* accum_noise_dev, size_t zk_rows) {
  CTX_GUARD();
  B2Z_REQUIRE(c, "null circuit");
  (void)code;
  return launch_accum(ctx, accum, data, c->accum_col_dev, c->accum_col.size(), mix_dev, accum_noise_dev, po2, zk_rows);
}
b2z_err b2z_witgen(b2z_ctx* ctx, const b2z_circuit* c, unsigned po2, uint32_t* code, uint32_t* data, uint32_t* globals_dev,
                   uint32_t* accum_noise_dev, uint64_t seed, size_t zk_rows) {
  CTX_GUARD();
  B2Z_REQUIRE(c, "null circuit");
  return witgen_launch(ctx, c, po2, code, data, globals_dev, accum_noise_dev, seed, zk_rows);
}

b2z_err b2z_prove_segment_device(b2z_ctx* ctx, const b2z_circuit* c, unsigned po2, const uint32_t* code_dev, const uint32_t* data_dev,
                                 const uint32_t* globals_host, const uint32_t* accum_noise_dev, size_t zk_rows, uint32_t* seal_host,
                                 size_t* seal_words) {
  CTX_GUARD();
  B2Z_REQUIRE(c && seal_host, "null argument");
  return prove_segment_device(ctx, c, po2, code_dev, data_dev, globals_host, accum_noise_dev, zk_rows, seal_host, seal_words, nullptr);
}
b2z_err b2z_prove_segment(b2z_ctx* ctx, const b2z_circuit* c, unsigned po2, const uint32_t* code_host, const uint32_t* data_host,
                          const uint32_t* globals_host, const uint32_t* accum_noise_host, size_t zk_rows, uint32_t* seal_host,
                          size_t* seal_words) {
  CTX_GUARD();
  B2Z_REQUIRE(c && seal_host, "null argument");
  return prove_segment_host(ctx, c, po2, code_host, data_host, globals_host, accum_noise_host, zk_rows, seal_host, seal_words);
}
b2z_err b2z_compute_control_root(b2z_ctx* ctx, unsigned po2, const uint32_t* code_dev, size_t code_cols, uint32_t* root8_host) {
  CTX_GUARD();
  B2Z_REQUIRE(code_dev && root8_host, "null argument");
  return compute_control_root(ctx, po2, code_dev, code_cols, root8_host);
}
b2z_err b2z_circuit_set_control_root(b2z_circuit* c, unsigned po2, const uint32_t* root8) {
  B2Z_REQUIRE(c && root8, "null argument");
  B2Z_REQUIRE(po2 >= 1 && po2 < 32, "po2 out of range");
  memcpy(c->control_root[po2], root8, 32);
  c->has_control_root[po2] = true;
  return nullptr;
}
int b2z_circuit_get_control_root(const b2z_circuit* c, unsigned po2, uint32_t* root8) {
  if (!c || po2 >= 32 || !c->has_control_root[po2]) return 0;
  if (root8) memcpy(root8, c->control_root[po2], 32);
  return 1;
}
size_t b2z_last_phase_ms(const b2z_ctx* ctx, double* ms, size_t cap) {
  if (!ctx) return 0;
  size_t n = ctx->phase_ms.size() < cap ? ctx->phase_ms.size() : cap;
  for (size_t i = 0; i < n; i++) ms[i] = ctx->phase_ms[i];
  return ctx->phase_ms.size();
}
const char* b2z_phase_name(size_t i) { return phase_name(i); }

int b2z_verify_segment(const b2z_circuit* c, const uint32_t* seal, size_t seal_words) {
  if (!c || !seal) return -1;
  return verify_segment(c, seal, seal_words);
}
void b2z_host_poseidon2_mix(uint32_t* s) { host_poseidon2_mix(s); }
void b2z_host_poseidon2_mix_f64(uint32_t* s) { host_poseidon2_mix_f64(s); }
void b2z_host_hash_elem_slice(const uint32_t* e, size_t n, uint32_t* d) { host_hash_elem_slice(e, n, d); }

}  // extern "C"
