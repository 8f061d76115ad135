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

namespace b2z {

__constant__ P2Consts c_p2;
static bool g_p2_uploaded[16] = {};

static b2z_err upload_consts(b2z_ctx* ctx) {
  static std::mutex m;
  std::lock_guard<std::mutex> lock(m);
  if (ctx->device < 16 && g_p2_uploaded[ctx->device]) return nullptr;
  const P2Consts& h = p2_host_consts();
  B2Z_CUDA(cudaMemcpyToSymbol(c_p2, &h, sizeof(P2Consts), 0, cudaMemcpyHostToDevice));
  if (ctx->device < 16) g_p2_uploaded[ctx->device] = true;
  return nullptr;
}

template <int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) hash_rows_kernel(uint32_t* __restrict__ out, const uint32_t* __restrict__ matrix, size_t rows,
                                                                  size_t cols, uint32_t ones) {
  size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (r >= rows) return;
  uint32_t s[P2_CELLS];
#pragma unroll
  for (int i = 0; i < P2_CELLS; i++) s[i] = 0;
  const uint32_t* p = matrix + r;
  size_t full = cols / P2_RATE;
  for (size_t blk = 0; blk < full; blk++) {
#pragma unroll
    for (int i = 0; i < P2_RATE; i++) s[i] = __ldg(p + (blk * P2_RATE + i) * rows);
    p2_permute_dev(s, c_p2, ones);
  }
  size_t rem = cols - full * P2_RATE;
  if (rem != 0 || cols == 0) {
#pragma unroll
    for (int i = 0; i < P2_RATE; i++) s[i] = (size_t)i < rem ? __ldg(p + (full * P2_RATE + i) * rows) : 0u;
    p2_permute_dev(s, c_p2, ones);
  }
  uint4* o = reinterpret_cast<uint4*>(out + 8 * r);
  o[0] = make_uint4(s[0], s[1], s[2], s[3]);
  o[1] = make_uint4(s[4], s[5], s[6], s[7]);
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

b2z_err launch_hash_rows(b2z_ctx* ctx, uint32_t* out, const uint32_t* matrix, size_t rows, size_t cols) {
  if (rows == 0) return nullptr;
  B2Z_TRY(upload_consts(ctx));
  // launch-shape experiments: B2Z_HASH_VARIANT = 0 (128 threads, compiler-chosen registers) | 1 (<= 48 regs) | 2 (<= 40 regs) | 3 (256 threads)
  static int variant = [] { const char* e = getenv("B2Z_HASH_VARIANT"); return e ? atoi(e) : 0; }();
  switch (variant) {
    case 1: hash_rows_kernel<128, 10><<<(unsigned)((rows + 127) / 128), 128, 0, ctx->stream>>>(out, matrix, rows, cols, 0xffffffffu); break;
    case 2: hash_rows_kernel<128, 12><<<(unsigned)((rows + 127) / 128), 128, 0, ctx->stream>>>(out, matrix, rows, cols, 0xffffffffu); break;
    case 3: hash_rows_kernel<256, 4><<<(unsigned)((rows + 255) / 256), 256, 0, ctx->stream>>>(out, matrix, rows, cols, 0xffffffffu); break;
    case 4: hash_rows_kernel<64, 16><<<(unsigned)((rows + 63) / 64), 64, 0, ctx->stream>>>(out, matrix, rows, cols, 0xffffffffu); break;
    default: hash_rows_kernel<128, 1><<<(unsigned)((rows + 127) / 128), 128, 0, ctx->stream>>>(out, matrix, rows, cols, 0xffffffffu); break;
  }
  ctx->launches++;
  B2Z_CUDA(cudaGetLastError());
  return nullptr;
}

b2z_err launch_hash_fold(b2z_ctx* ctx, uint32_t* nodes, size_t in_size, size_t out_size) {
  if (out_size == 0) return nullptr;
  B2Z_TRY(upload_consts(ctx));
  hash_fold_kernel<<<(unsigned)((out_size + 127) / 128), 128, 0, ctx->stream>>>(nodes, in_size, out_size, 0xffffffffu);
  ctx->launches++;
  B2Z_CUDA(cudaGetLastError());
  return nullptr;
}

}  // namespace b2z
