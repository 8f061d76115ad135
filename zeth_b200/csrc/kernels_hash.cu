// Poseidon2 Merkle kernels: hash_rows (leaf sponge over a column-major matrix) and hash_fold (hash_pair layer).
//
// Semantics: risc0-zkp 3.0.2 hal/cpu.rs hash_rows / hash_fold with the Poseidon2 suite
// (core/hash/poseidon2/mod.rs unpadded_hash: overwrite-mode sponge, rate 16, zero-padded tail; SURVEY.md A.4/A.5,
// scope rows a9/a10).  One thread owns one row (or one parent node): the 24-word state stays in registers, loads
// are coalesced across the warp because consecutive threads read consecutive rows of the same column.
// These kernels are INT32-ALU bound (1 356 modular multiplies per 64 bytes absorbed), not HBM bound.
#include <mutex>

#include "common.h"
#include "poseidon2.cuh"
#include "poseidon2_consts.h"
#include "poseidon2_f64.cuh"

namespace b2z {

__constant__ P2Consts c_p2;
__constant__ P2ConstsF64 c_p2f;
static bool g_p2_uploaded[16] = {};

static b2z_err upload_consts(b2z_ctx* ctx) {
  static std::mutex m;
  std::lock_guard<std::mutex> lock(m);
  if (ctx->device < 16 && g_p2_uploaded[ctx->device]) return nullptr;
  const P2Consts& h = p2_host_consts();
  B2Z_CUDA(cudaMemcpyToSymbol(c_p2, &h, sizeof(P2Consts), 0, cudaMemcpyHostToDevice));
  const P2ConstsF64& hf = p2_host_consts_f64();
  B2Z_CUDA(cudaMemcpyToSymbol(c_p2f, &hf, sizeof(P2ConstsF64), 0, cudaMemcpyHostToDevice));
  if (ctx->device < 16) g_p2_uploaded[ctx->device] = true;
  return nullptr;
}

// Column-chunked absorption (used when a witness is still arriving over PCIe, prover.cpp commit_group): a launch may
// start from the capacity cells left by the previous chunk (HASH_CARRY_IN) and may stop before squeezing (no HASH_LAST).
// The 8 carried words of a row live in that row's digest slot, as canonical Montgomery words, so either stream can
// resume a row the other one started.  In overwrite mode the rate cells never need saving.  Non-final chunks must be
// whole blocks (cols % 16 == 0, cols > 0).
constexpr unsigned HASH_CARRY_IN = 1, HASH_LAST = 2;

template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) hash_rows_kernel(uint32_t* __restrict__ out, const uint32_t* __restrict__ matrix, size_t rows,
                                                                  size_t cols, uint32_t ones, unsigned flags) {
  size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (r >= rows) return;
  uint32_t s[P2_CELLS];
#pragma unroll
  for (int i = 0; i < P2_CELLS; i++) s[i] = 0;
  uint4* o = reinterpret_cast<uint4*>(out + 8 * r);
  if (flags & HASH_CARRY_IN) {
    uint4 a = o[0], b = o[1];
    s[16] = a.x, s[17] = a.y, s[18] = a.z, s[19] = a.w, s[20] = b.x, s[21] = b.y, s[22] = b.z, s[23] = b.w;
  }
  const uint32_t* p = matrix + r;
  size_t full = cols / P2_RATE;
  for (size_t blk = 0; blk < full; blk++) {
#pragma unroll
    for (int i = 0; i < P2_RATE; i++) s[i] = __ldg(p + (blk * P2_RATE + i) * rows);
    p2_permute_dev(s, c_p2, ones);
  }
  size_t rem = cols - full * P2_RATE;
  if (rem != 0 || (cols == 0 && !(flags & HASH_CARRY_IN))) {
#pragma unroll
    for (int i = 0; i < P2_RATE; i++) s[i] = (size_t)i < rem ? __ldg(p + (full * P2_RATE + i) * rows) : 0u;
    p2_permute_dev(s, c_p2, ones);
  }
  if (flags & HASH_LAST) {
    o[0] = make_uint4(s[0], s[1], s[2], s[3]);
    o[1] = make_uint4(s[4], s[5], s[6], s[7]);
  } else {
    o[0] = make_uint4(s[16], s[17], s[18], s[19]);
    o[1] = make_uint4(s[20], s[21], s[22], s[23]);
  }
}

// ---- mixed INT / FP64 leaf hashing ----
// One persistent launch; every warp pulls 32-row tiles from a global counter until none are left.  A CTA is either
// an INT CTA (poseidon2.cuh: IMAD + ALU pipes) or an FP64 CTA (poseidon2_f64.cuh: DFMA pipe only); the role comes
// from a per-SM arrival counter so that every SM hosts both kinds in the ratio n_int : (period - n_int).  The two
// instruction streams produce identical digests, so which warp hashed a row is unobservable; the dynamic tile queue
// balances their different speeds.  sched[0] = tile counter, sched[1 + smid] = CTA arrival counter of that SM.
constexpr int HASH_SCHED_WORDS = 1 + 512;

__device__ __forceinline__ double p2f_reduce(double x) { return p2f_red(x); }  // |x| < 2^38 -> (-0.51 P, 0.51 P)

// both row routines keep a single call site of their permutation (tail columns are predicated loads of 0).
// The guarded load is spelled in PTX: left to itself the compiler widens the guard to 64 bits and rebuilds every
// address from scratch (9 instructions per column) -- code size matters here, see the instruction-cache note above.
__device__ __forceinline__ uint32_t ld_col_guarded(const uint32_t* p, unsigned i, unsigned left) {
  uint32_t v;
  asm volatile(
      "{\n\t.reg .pred q;\n\tsetp.lt.u32 q, %2, %3;\n\tmov.u32 %0, 0;\n\t@q ld.global.nc.u32 %0, [%1];\n\t}"
      : "=r"(v)
      : "l"(p), "r"(i), "r"(left));
  return v;
}
__device__ __forceinline__ void hash_row_int(uint32_t* __restrict__ out, const uint32_t* __restrict__ matrix, size_t rows, size_t cols, size_t r,
                                             uint32_t ones, unsigned flags) {
  uint32_t s[P2_CELLS];
#pragma unroll
  for (int i = 0; i < P2_CELLS; i++) s[i] = 0;
  uint4* o = reinterpret_cast<uint4*>(out + 8 * r);
  if (flags & HASH_CARRY_IN) {
    uint4 a = o[0], b = o[1];
    s[16] = a.x, s[17] = a.y, s[18] = a.z, s[19] = a.w, s[20] = b.x, s[21] = b.y, s[22] = b.z, s[23] = b.w;
  }
  const uint32_t* p = matrix + r;
  const size_t nblk = (cols == 0 && !(flags & HASH_CARRY_IN)) ? 1 : (cols + P2_RATE - 1) / P2_RATE;
#pragma unroll 1
  for (size_t blk = 0; blk < nblk; blk++) {
    const unsigned left = (unsigned)min(cols - blk * P2_RATE, (size_t)P2_RATE);  // columns in this block (0 only when cols == 0)
#pragma unroll
    for (int i = 0; i < P2_RATE; i++) {
      s[i] = ld_col_guarded(p, (unsigned)i, left);
      p += rows;
    }
    p2_permute_dev_compact(s, c_p2, ones);
  }
  const int base = (flags & HASH_LAST) ? 0 : P2_RATE;  // squeeze the digest, or park the capacity cells
  uint32_t w[8];
#pragma unroll
  for (int i = 0; i < 8; i++) w[i] = base ? s[P2_RATE + i] : s[i];
  o[0] = make_uint4(w[0], w[1], w[2], w[3]);
  o[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

__device__ __forceinline__ void hash_row_f64(uint32_t* __restrict__ out, const uint32_t* __restrict__ matrix, size_t rows, size_t cols, size_t r,
                                             unsigned flags) {
  double s[P2_CELLS];
#pragma unroll
  for (int i = 0; i < P2_CELLS; i++) s[i] = 0.0;
  uint4* o = reinterpret_cast<uint4*>(out + 8 * r);
  if (flags & HASH_CARRY_IN) {
    uint4 a = o[0], b = o[1];
    uint32_t w[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 8; i++) s[P2_RATE + i] = p2f_from_mont(w[i]);
  }
  const uint32_t* p = matrix + r;
  const size_t nblk = (cols == 0 && !(flags & HASH_CARRY_IN)) ? 1 : (cols + P2_RATE - 1) / P2_RATE;
#pragma unroll 1
  for (size_t blk = 0; blk < nblk; blk++) {
    const unsigned left = (unsigned)min(cols - blk * P2_RATE, (size_t)P2_RATE);
#pragma unroll
    for (int i = 0; i < P2_RATE; i++) {
      s[i] = p2f_from_mont(ld_col_guarded(p, (unsigned)i, left));
      p += rows;
    }
#pragma unroll
    for (int i = P2_RATE; i < P2_CELLS; i++) s[i] = p2f_reduce(s[i]);  // carried capacity cells back below 2^31
    p2f_permute(s, c_p2f);
  }
  const int base = (flags & HASH_LAST) ? 0 : P2_RATE;
  uint32_t w[8];
#pragma unroll
  for (int i = 0; i < 8; i++) w[i] = p2f_to_mont(base ? s[P2_RATE + i] : s[i]);
  o[0] = make_uint4(w[0], w[1], w[2], w[3]);
  o[1] = make_uint4(w[4], w[5], w[6], w[7]);
}

template <int MINB>
__global__ void __launch_bounds__(128, MINB) hash_rows_mixed_kernel(uint32_t* __restrict__ out, const uint32_t* __restrict__ matrix, size_t rows,
                                                                    size_t cols, uint32_t ones, unsigned* __restrict__ sched, unsigned n_tiles,
                                                                    unsigned period, unsigned n_int, unsigned flags) {
  __shared__ unsigned s_role;
  if (threadIdx.x == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    s_role = atomicAdd(&sched[1 + (smid & 511)], 1u) % period;
  }
  __syncthreads();
  const bool f64_role = s_role >= n_int;
  const unsigned lane = threadIdx.x & 31;
  for (;;) {
    unsigned tile = 0;
    if (lane == 0) tile = atomicAdd(&sched[0], 1u);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    if (tile >= n_tiles) break;
    size_t r = (size_t)tile * 32 + lane;
    if (r < rows) {
      if (f64_role) hash_row_f64(out, matrix, rows, cols, r, flags);
      else hash_row_int(out, matrix, rows, cols, r, ones, flags);
    }
  }
}

// Role-specialised persistent kernels (B2Z_HASH_SPLIT=1): the same tile queue, but the INT stream and the FP64 stream are
// two LAUNCHES on two streams.  A mixed kernel is compiled for the larger of the two bodies (80 registers -> 6 CTAs = 24 warps per
// SM); apart, the INT body needs 48 registers and the FP64 body 80, so an SM can host e.g. 4 + 4 CTAs (32 warps) -- more
// warps to cover the dependent product chains that keep the mixed kernel's issue rate at ~69 %.
template <int MINB>
__global__ void __launch_bounds__(128, MINB) hash_rows_int_kernel(uint32_t* __restrict__ out, const uint32_t* __restrict__ matrix, size_t rows,
                                                                  size_t cols, uint32_t ones, unsigned* __restrict__ sched, unsigned n_tiles,
                                                                  unsigned flags) {
  const unsigned lane = threadIdx.x & 31;
  for (;;) {
    unsigned tile = 0;
    if (lane == 0) tile = atomicAdd(&sched[0], 1u);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    if (tile >= n_tiles) break;
    size_t r = (size_t)tile * 32 + lane;
    if (r < rows) hash_row_int(out, matrix, rows, cols, r, ones, flags);
  }
}
template <int MINB>
__global__ void __launch_bounds__(128, MINB) hash_rows_f64_kernel(uint32_t* __restrict__ out, const uint32_t* __restrict__ matrix, size_t rows,
                                                                  size_t cols, unsigned* __restrict__ sched, unsigned n_tiles, unsigned flags) {
  const unsigned lane = threadIdx.x & 31;
  for (;;) {
    unsigned tile = 0;
    if (lane == 0) tile = atomicAdd(&sched[0], 1u);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    if (tile >= n_tiles) break;
    size_t r = (size_t)tile * 32 + lane;
    if (r < rows) hash_row_f64(out, matrix, rows, cols, r, flags);
  }
}

__global__ void __launch_bounds__(128) hash_fold_kernel(uint32_t* __restrict__ nodes, size_t in_size, size_t out_size, uint32_t ones) {
  size_t j = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (j >= out_size) return;
  const uint4* in = reinterpret_cast<const uint4*>(nodes + 8 * (in_size + 2 * j));
  uint4 a = in[0], b = in[1], c = in[2], d = in[3];
  uint32_t s[P2_CELLS] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w, 0, 0, 0, 0, 0, 0, 0, 0};
  p2_permute_dev(s, c_p2, ones);
  uint4* o = reinterpret_cast<uint4*>(nodes + 8 * (out_size + j));
  o[0] = make_uint4(s[0], s[1], s[2], s[3]);
  o[1] = make_uint4(s[4], s[5], s[6], s[7]);
}

// Merkle interior layer with the same two streams: a tile is 32 parents, one per lane.
template <int MINB>
__global__ void __launch_bounds__(128, MINB) hash_fold_mixed_kernel(uint32_t* __restrict__ nodes, size_t in_size, size_t out_size, uint32_t ones,
                                                                    unsigned* __restrict__ sched, unsigned n_tiles, unsigned period, unsigned n_int) {
  __shared__ unsigned s_role;
  if (threadIdx.x == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    s_role = atomicAdd(&sched[1 + (smid & 511)], 1u) % period;
  }
  __syncthreads();
  const bool f64_role = s_role >= n_int;
  const unsigned lane = threadIdx.x & 31;
  for (;;) {
    unsigned tile = 0;
    if (lane == 0) tile = atomicAdd(&sched[0], 1u);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    if (tile >= n_tiles) break;
    size_t j = (size_t)tile * 32 + lane;
    if (j >= out_size) continue;
    const uint4* in = reinterpret_cast<const uint4*>(nodes + 8 * (in_size + 2 * j));
    uint4 a = in[0], b = in[1], c = in[2], d = in[3];
    uint32_t w[P2_RATE] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w};
    uint4 o0, o1;
    if (f64_role) {
      double s[P2_CELLS];
#pragma unroll
      for (int i = 0; i < P2_RATE; i++) s[i] = p2f_from_mont(w[i]);
#pragma unroll
      for (int i = P2_RATE; i < P2_CELLS; i++) s[i] = 0.0;
      p2f_permute(s, c_p2f);
      o0 = make_uint4(p2f_to_mont(s[0]), p2f_to_mont(s[1]), p2f_to_mont(s[2]), p2f_to_mont(s[3]));
      o1 = make_uint4(p2f_to_mont(s[4]), p2f_to_mont(s[5]), p2f_to_mont(s[6]), p2f_to_mont(s[7]));
    } else {
      uint32_t s[P2_CELLS];
#pragma unroll
      for (int i = 0; i < P2_RATE; i++) s[i] = w[i];
#pragma unroll
      for (int i = P2_RATE; i < P2_CELLS; i++) s[i] = 0u;
      p2_permute_dev_compact(s, c_p2, ones);
      o0 = make_uint4(s[0], s[1], s[2], s[3]);
      o1 = make_uint4(s[4], s[5], s[6], s[7]);
    }
    uint4* o = reinterpret_cast<uint4*>(nodes + 8 * (out_size + j));
    o[0] = o0;
    o[1] = o1;
  }
}

// Top of the tree in one launch: all layers from `start_out` parents down to the root are folded by a single CTA
// (layer after layer, __syncthreads between them) -- replaces ~10 tiny launches per Merkle tree.
constexpr unsigned FOLD_TOP_THREADS = 512;
__global__ void __launch_bounds__(FOLD_TOP_THREADS) hash_fold_top_kernel(uint32_t* __restrict__ nodes, unsigned start_out, uint32_t ones) {
  for (unsigned out_size = start_out; out_size >= 1; out_size >>= 1) {
    for (unsigned j = threadIdx.x; j < out_size; j += blockDim.x) {
      const uint4* in = reinterpret_cast<const uint4*>(nodes + 8 * (size_t)(2 * out_size + 2 * j));
      uint4 a = in[0], b = in[1], c = in[2], d = in[3];
      uint32_t s[P2_CELLS] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w, d.x, d.y, d.z, d.w, 0, 0, 0, 0, 0, 0, 0, 0};
      p2_permute_dev(s, c_p2, ones);
      uint4* o = reinterpret_cast<uint4*>(nodes + 8 * (size_t)(out_size + j));
      o[0] = make_uint4(s[0], s[1], s[2], s[3]);
      o[1] = make_uint4(s[4], s[5], s[6], s[7]);
    }
    __syncthreads();
  }
}
b2z_err launch_hash_fold_top(b2z_ctx* ctx, uint32_t* nodes, size_t start_out) {
  if (start_out == 0) return nullptr;
  B2Z_TRY(upload_consts(ctx));
  hash_fold_top_kernel<<<1, FOLD_TOP_THREADS, 0, ctx->stream>>>(nodes, (unsigned)start_out, 0xffffffffu);
  ctx->launches++;
  B2Z_CUDA(cudaGetLastError());
  return nullptr;
}

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

b2z_err launch_hash_rows(b2z_ctx* ctx, uint32_t* out, const uint32_t* matrix, size_t rows, size_t cols) {
  return launch_hash_rows_chunk(ctx, out, matrix, rows, cols, false, true);
}

b2z_err launch_hash_rows_chunk(b2z_ctx* ctx, uint32_t* out, const uint32_t* matrix, size_t rows, size_t cols, bool carry_in, bool last) {
  if (rows == 0) return nullptr;
  B2Z_REQUIRE(last || (cols > 0 && cols % P2_RATE == 0), "hash_rows_chunk: a non-final chunk must be a positive multiple of 16 columns");
  // overwrite-mode sponge: only the 8 capacity words are carried between chunks, so an EMPTY final chunk has nothing to
  // squeeze from (the rate cells of the previous permutation are gone) -- reject instead of writing zero digests
  B2Z_REQUIRE(!(carry_in && last && cols == 0), "hash_rows_chunk: the final chunk of a carried sponge must not be empty");
  const unsigned flags = (carry_in ? HASH_CARRY_IN : 0u) | (last ? HASH_LAST : 0u);
  B2Z_TRY(upload_consts(ctx));
  // B2Z_HASH_MIXED=0 forces the INT-only kernel; the mixed kernel needs enough rows to fill the persistent grid
  // (read per call so that tests and sweeps can switch modes inside one process)
  const int mixed = env_int("B2Z_HASH_MIXED", 1);
  const int ctas_per_sm = env_int("B2Z_HASH_CTAS", 6);          // resident CTAs per SM: 6 (80 regs), 7 (72), 8 (64, spills)
  const int period = env_int("B2Z_HASH_ROLE_PERIOD", 2);        // of every `period` CTAs on an SM ...
  const int n_int = env_int("B2Z_HASH_ROLE_INT", 1);            // ... this many run the INT stream, the rest FP64
  const size_t grid = (size_t)ctx->sm_count * ctas_per_sm;
  const size_t min_rows = (size_t)env_int("B2Z_HASH_MIXED_MIN_ROWS", (int)(grid * 128 * 4));
  const int split = env_int("B2Z_HASH_SPLIT", 0);
  if (mixed && split && rows >= min_rows && rows / 32 < 0xffffffffu) {
    // two role-specialised launches sharing one tile queue; the FP64 launch runs on the auxiliary stream (fork / join by events)
    const int ci = env_int("B2Z_HASH_SPLIT_INT", 4), cf = env_int("B2Z_HASH_SPLIT_F64", 4);
    if (!ctx->hash_sched) B2Z_CUDA(cudaMalloc(&ctx->hash_sched, HASH_SCHED_WORDS * sizeof(unsigned)));
    if (!ctx->aux_stream) {
      B2Z_CUDA(cudaStreamCreateWithFlags(&ctx->aux_stream, cudaStreamNonBlocking));
      B2Z_CUDA(cudaEventCreateWithFlags(&ctx->aux_fork, cudaEventDisableTiming));
      B2Z_CUDA(cudaEventCreateWithFlags(&ctx->aux_join, cudaEventDisableTiming));
    }
    B2Z_CUDA(cudaMemsetAsync(ctx->hash_sched, 0, HASH_SCHED_WORDS * sizeof(unsigned), ctx->stream));
    unsigned n_tiles = (unsigned)((rows + 31) / 32);
    B2Z_CUDA(cudaEventRecord(ctx->aux_fork, ctx->stream));
    B2Z_CUDA(cudaStreamWaitEvent(ctx->aux_stream, ctx->aux_fork, 0));
    if (ci > 0)
      hash_rows_int_kernel<10><<<(unsigned)(ctx->sm_count * ci), 128, 0, ctx->stream>>>(out, matrix, rows, cols, 0xffffffffu, ctx->hash_sched, n_tiles, flags);
    if (cf > 0)
      hash_rows_f64_kernel<6><<<(unsigned)(ctx->sm_count * cf), 128, 0, ctx->aux_stream>>>(out, matrix, rows, cols, ctx->hash_sched, n_tiles, flags);
    B2Z_CUDA(cudaEventRecord(ctx->aux_join, ctx->aux_stream));
    B2Z_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->aux_join, 0));
    ctx->launches++;
  } else if (mixed && period >= 1 && rows >= min_rows && rows / 32 < 0xffffffffu) {
    if (!ctx->hash_sched) B2Z_CUDA(cudaMalloc(&ctx->hash_sched, HASH_SCHED_WORDS * sizeof(unsigned)));
    B2Z_CUDA(cudaMemsetAsync(ctx->hash_sched, 0, HASH_SCHED_WORDS * sizeof(unsigned), ctx->stream));
    unsigned n_tiles = (unsigned)((rows + 31) / 32);
    auto kern = ctas_per_sm >= 8 ? hash_rows_mixed_kernel<8> : ctas_per_sm == 7 ? hash_rows_mixed_kernel<7> : hash_rows_mixed_kernel<6>;
    kern<<<(unsigned)grid, 128, 0, ctx->stream>>>(out, matrix, rows, cols, 0xffffffffu, ctx->hash_sched, n_tiles, (unsigned)period,
                                                  (unsigned)n_int, flags);
  } else {
    hash_rows_kernel<128, 1><<<(unsigned)((rows + 127) / 128), 128, 0, ctx->stream>>>(out, matrix, rows, cols, 0xffffffffu, flags);
  }
  ctx->launches++;
  B2Z_CUDA(cudaGetLastError());
  return nullptr;
}

b2z_err launch_hash_fold(b2z_ctx* ctx, uint32_t* nodes, size_t in_size, size_t out_size) {
  if (out_size == 0) return nullptr;
  B2Z_TRY(upload_consts(ctx));
  const int mixed = env_int("B2Z_HASH_MIXED", 1);
  const int ctas_per_sm = env_int("B2Z_HASH_CTAS", 6);
  const int period = env_int("B2Z_HASH_ROLE_PERIOD", 2);
  const int n_int = env_int("B2Z_HASH_ROLE_INT", 1);
  const size_t grid = (size_t)ctx->sm_count * ctas_per_sm;
  const size_t min_rows = (size_t)env_int("B2Z_HASH_MIXED_MIN_ROWS", (int)(grid * 128 * 4));
  if (mixed && period >= 1 && out_size >= min_rows && out_size / 32 < 0xffffffffu) {
    if (!ctx->hash_sched) B2Z_CUDA(cudaMalloc(&ctx->hash_sched, HASH_SCHED_WORDS * sizeof(unsigned)));
    B2Z_CUDA(cudaMemsetAsync(ctx->hash_sched, 0, HASH_SCHED_WORDS * sizeof(unsigned), ctx->stream));
    unsigned n_tiles = (unsigned)((out_size + 31) / 32);
    auto kern = ctas_per_sm >= 8 ? hash_fold_mixed_kernel<8> : ctas_per_sm == 7 ? hash_fold_mixed_kernel<7> : hash_fold_mixed_kernel<6>;
    kern<<<(unsigned)grid, 128, 0, ctx->stream>>>(nodes, in_size, out_size, 0xffffffffu, ctx->hash_sched, n_tiles, (unsigned)period, (unsigned)n_int);
  } else {
    hash_fold_kernel<<<(unsigned)((out_size + 127) / 128), 128, 0, ctx->stream>>>(nodes, in_size, out_size, 0xffffffffu);
  }
  ctx->launches++;
  B2Z_CUDA(cudaGetLastError());
  return nullptr;
}

}  // namespace b2z
