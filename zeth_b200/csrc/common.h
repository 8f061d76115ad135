// Shared host-side declarations of the product library (context, error plumbing, kernel launch entry points).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <cstdlib>
#include <vector>

#include "../../include/b2z.h"
#include "field.cuh"

namespace b2z {

// thread-local error slot: entry points return its c_str() (upstream FFI convention: NULL == success)
const char* set_error(const std::string& msg);

#define B2Z_CUDA(expr)                                                                              \
  do {                                                                                              \
    cudaError_t _e = (expr);                                                                        \
    if (_e != cudaSuccess)                                                                          \
      return ::b2z::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e));                  \
  } while (0)
#define B2Z_TRY(expr)                  \
  do {                                 \
    const char* _err = (expr);         \
    if (_err) return _err;             \
  } while (0)
#define B2Z_REQUIRE(cond, msg)                       \
  do {                                               \
    if (!(cond)) return ::b2z::set_error(msg);       \
  } while (0)

constexpr int MAX_PO2 = 26;  // largest transform: po2 24 trace x 4 (INV_RATE)

}  // namespace b2z

struct b2z_ctx {
  int device = 0;
  int sm_count = 148;
  cudaStream_t own_stream = nullptr;
  cudaStream_t stream = nullptr;
  cudaStream_t copy_stream = nullptr;          // host->device witness uploads overlapped with compute (b2z_prove_segment)
  std::vector<cudaEvent_t> h2d_events;
  std::vector<cudaEvent_t> event_pool;         // recycled timing-disabled events of the witness uploads
  // pinned staging ring for PAGEABLE host witnesses (b2z_prove_segment): slots of 16 MB, one drain event per slot
  char* stage_ring = nullptr;
  std::vector<cudaEvent_t> stage_events;
  bool last_upload_staged = false;
  uint64_t launches = 0;
  cudaEvent_t wait_event = nullptr;            // cudaEventBlockingSync event behind b2z::stream_wait
  // per-size device table, built lazily: 3^rev(j) for zk_shift (NTT twist tables live in kernels_ntt.cu's registry)
  uint32_t* pow3rev[b2z::MAX_PO2 + 1] = {};
  cudaStream_t aux_stream = nullptr;           // second launch stream of the role-split hash kernels (fork / join by events)
  cudaEvent_t aux_fork = nullptr, aux_join = nullptr;
  unsigned* hash_sched = nullptr;  // tile queue + per-SM role counters of the mixed INT/FP64 hash kernels
  // pinned staging for small transcript round trips
  uint32_t* pinned = nullptr;
  size_t pinned_words = 0;
  // phase timing of the last prove call
  std::vector<double> phase_ms;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> phase_events;
};

namespace b2z {

// Host wait for everything enqueued on `s`: an OS-blocking wait on a cudaEventBlockingSync event instead of
// cudaStreamSynchronize's spin loop.  A session runs `in_flight` worker threads per GPU and a node runs one process per
// GPU: spinning waits would burn 2 x 8 cores (and, inside a CPU-quota cgroup, get every thread of the process throttled).
// B2Z_SPIN_SYNC=1 restores the spinning wait.
inline cudaError_t stream_wait(b2z_ctx* ctx, cudaStream_t s) {
  static const bool spin = [] { const char* e = getenv("B2Z_SPIN_SYNC"); return e && e[0] == '1'; }();
  if (spin) return cudaStreamSynchronize(s);
  if (!ctx->wait_event) {
    cudaError_t e = cudaEventCreateWithFlags(&ctx->wait_event, cudaEventBlockingSync | cudaEventDisableTiming);
    if (e != cudaSuccess) return e;
  }
  cudaError_t e = cudaEventRecord(ctx->wait_event, s);
  if (e != cudaSuccess) return e;
  return cudaEventSynchronize(ctx->wait_event);
}

// ---- kernel launchers (kernels_*.cu); all enqueue on ctx->stream and bump ctx->launches ----
b2z_err ensure_tables(b2z_ctx* ctx, unsigned po2, bool pow3);
// `src` (optional): read the natural-order evaluations from a separate buffer (saves the witness -> coefficient copy)
b2z_err launch_interpolate(b2z_ctx* ctx, uint32_t* io, size_t count, unsigned po2, bool fuse_zk_shift, const uint32_t* src = nullptr);
b2z_err launch_evaluate(b2z_ctx* ctx, uint32_t* out, const uint32_t* in, size_t count, unsigned out_po2, unsigned expand_bits);
b2z_err launch_bit_reverse(b2z_ctx* ctx, uint32_t* io, size_t count, unsigned po2);
b2z_err launch_zk_shift(b2z_ctx* ctx, uint32_t* io, size_t count, unsigned po2);
b2z_err launch_hash_rows(b2z_ctx* ctx, uint32_t* out, const uint32_t* matrix, size_t rows, size_t cols);
// column-chunked leaf hashing: `out` (8 words per row) carries the sponge capacity between chunks, the digest after the last
b2z_err launch_hash_rows_chunk(b2z_ctx* ctx, uint32_t* out, const uint32_t* matrix, size_t rows, size_t cols, bool carry_in, bool last);
b2z_err launch_hash_fold(b2z_ctx* ctx, uint32_t* nodes, size_t in_size, size_t out_size);
// all fold layers with <= start_out parents, one launch (node layout as hash_fold: in_size = 2 * out_size)
b2z_err launch_hash_fold_top(b2z_ctx* ctx, uint32_t* nodes, size_t start_out);
b2z_err launch_fri_fold(b2z_ctx* ctx, uint32_t* out, const uint32_t* in, const Fp4& mix, size_t out_count);
b2z_err launch_evaluate_any(b2z_ctx* ctx, const uint32_t* coeffs, unsigned po2, const uint32_t* which, const uint32_t* xs,
                            size_t eval_count, uint32_t* out);
b2z_err launch_evaluate_any_host(b2z_ctx* ctx, const uint32_t* coeffs, unsigned po2, const uint32_t* which_host, const Fp4* xs_host,
                                 size_t eval_count, uint32_t* out);
b2z_err launch_mix_poly_coeffs(b2z_ctx* ctx, uint32_t* out, const Fp4& mix_start, const Fp4& mix, const uint32_t* in,
                               const uint32_t* combos_dev, size_t input_size, size_t count);
b2z_err launch_mix_poly_coeffs_host(b2z_ctx* ctx, uint32_t* out, const Fp4& mix_start, const Fp4& mix, const uint32_t* in,
                                    const uint32_t* combos_host, size_t input_size, size_t count);
b2z_err launch_add_elem(b2z_ctx* ctx, uint32_t* out, const uint32_t* a, const uint32_t* b, size_t n);
b2z_err launch_sum_extelem(b2z_ctx* ctx, uint32_t* out, const uint32_t* in, size_t count, size_t to_add);
b2z_err launch_zeroize(b2z_ctx* ctx, uint32_t* io, size_t n);
b2z_err launch_gather_sample(b2z_ctx* ctx, uint32_t* dst, const uint32_t* src, size_t idx, size_t size, size_t stride);
b2z_err launch_poly_divide(b2z_ctx* ctx, uint32_t* ext_coeffs, size_t count, const Fp4& z, uint32_t* remainder_dev);
b2z_err launch_sub_lowcoeffs(b2z_ctx* ctx, uint32_t* ext_coeffs, const uint32_t* sub_dev, size_t n_sub);
b2z_err launch_accum(b2z_ctx* ctx, uint32_t* accum, const uint32_t* data, const uint32_t* accum_cols_dev, size_t n_accum,
                     const uint32_t* mix_dev, const uint32_t* noise_dev, unsigned po2, size_t zk_rows);
b2z_err launch_scatter(b2z_ctx* ctx, uint32_t* into, const uint32_t* index, const uint32_t* offsets, const uint32_t* values, size_t count);
b2z_err launch_prefix_products(b2z_ctx* ctx, uint32_t* ext_io, size_t n);
// query gathering: writes rows + Merkle paths for all queries straight into the device seal image
struct GatherTree {
  const uint32_t* matrix;
  const uint32_t* nodes;
  uint32_t rows, cols, top_size;
  uint32_t seal_off;  // word offset of this tree's section inside one query record; row = pos & (rows - 1)
};
b2z_err launch_gather_queries(b2z_ctx* ctx, uint32_t* seal_dev, const GatherTree* trees_dev, size_t n_trees,
                              const uint32_t* positions_dev, size_t n_queries, size_t record_words);

// host Poseidon2 / transcript (transcript.cpp)
void host_poseidon2_mix(uint32_t* state24);
void host_poseidon2_mix_f64(uint32_t* state24);  // same permutation through the FP64 instruction stream
void host_hash_elem_slice(const uint32_t* in, size_t n, uint32_t* digest);
void host_hash_pair(const uint32_t* a, const uint32_t* b, uint32_t* digest);
struct HostRng {
  uint32_t cells[24] = {};
  unsigned pool_used = 0;
  void mix(const uint32_t* digest);
  uint32_t random_elem();
  Fp4 random_ext();
  uint32_t random_bits(unsigned bits);
};

}  // namespace b2z
