// BabyBear NTT kernels: batch_interpolate_ntt (DIF, natural -> bit-reversed, optional fused zk_shift),
// batch_expand_into_evaluate_ntt / batch_evaluate_ntt (DIT, bit-reversed -> natural), bit_reverse, zk_shift.
//
// Semantics follow risc0-zkp 3.0.2 core/ntt.rs + hal/cpu.rs (scope rows a5-a8).  Data layout is
// column-major [count][n].  A transform of 2^logn points runs as at most two passes over HBM: each CTA stages a
// tile of 2^T elements in shared memory and performs up to T butterfly layers on it.  The tile of the "far"
// pass is [2^B strided rows] x [2^LB contiguous elements] so that every global access is a >= 64-byte run.
#include "common.h"

namespace b2z {

__global__ void build_twiddles_kernel(uint32_t* tw, uint32_t w, size_t half) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < half) tw[i] = fp_pow(w, i);
}
__global__ void build_pow3rev_kernel(uint32_t* tab, unsigned bits) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < ((size_t)1 << bits)) tab[i] = fp_pow(MONT_THREE, bit_rev((uint32_t)i, bits));
}

b2z_err ensure_tables(b2z_ctx* ctx, unsigned po2, bool fwd, bool rev, bool pow3) {
  B2Z_REQUIRE(po2 <= MAX_PO2, "transform size exceeds 2^26");
  size_t n = (size_t)1 << po2, half = n / 2;
  if (fwd && !ctx->tw_fwd[po2]) {
    B2Z_CUDA(cudaMalloc(&ctx->tw_fwd[po2], sizeof(uint32_t) * (half ? half : 1)));
    if (half) build_twiddles_kernel<<<(unsigned)((half + 255) / 256), 256, 0, ctx->stream>>>(ctx->tw_fwd[po2], rou_fwd(po2), half);
    ctx->launches++;
  }
  if (rev && !ctx->tw_rev[po2]) {
    B2Z_CUDA(cudaMalloc(&ctx->tw_rev[po2], sizeof(uint32_t) * (half ? half : 1)));
    if (half) build_twiddles_kernel<<<(unsigned)((half + 255) / 256), 256, 0, ctx->stream>>>(ctx->tw_rev[po2], rou_rev(po2), half);
    ctx->launches++;
  }
  if (pow3 && !ctx->pow3rev[po2]) {
    B2Z_CUDA(cudaMalloc(&ctx->pow3rev[po2], sizeof(uint32_t) * n));
    build_pow3rev_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->pow3rev[po2], po2);
    ctx->launches++;
  }
  B2Z_CUDA(cudaGetLastError());
  return nullptr;
}

// One pass over butterfly layers (s_lo + t) for t in [t_begin, B] of a 2^logn-point transform, every column.
//   tile element (mid, lo): global index = hi << (s_lo + B) | mid << s_lo | (lo_blk << LB | lo)
//   INVERSE = false: decimation in time   a' = a + w b, b' = a - w b, layers ascending
//   INVERSE = true : decimation in frequency a' = a + b, b' = (a - b) w, layers descending
// `ebits` > 0 (forward first pass only): the tile is loaded from `in` at index (gidx >> ebits) -- the expand step.
template <bool INVERSE>
__global__ void ntt_pass_kernel(uint32_t* __restrict__ out, const uint32_t* __restrict__ in, const uint32_t* __restrict__ tw,
                                const uint32_t* __restrict__ scale_tab, uint32_t scale, unsigned logn, unsigned s_lo, unsigned B,
                                unsigned LB, unsigned t_begin, unsigned ebits, int apply_scale) {
  extern __shared__ uint32_t sm[];
  const unsigned tile_log = B + LB;
  const unsigned tile = 1u << tile_log;
  const size_t n = (size_t)1 << logn;
  const unsigned lo_blocks_log = s_lo - LB;
  const size_t tile_id = blockIdx.x;
  const size_t hi = tile_id >> lo_blocks_log;
  const unsigned lo_blk = (unsigned)(tile_id & ((1u << lo_blocks_log) - 1));
  const size_t col = blockIdx.y;
  const size_t base = (hi << (s_lo + B)) | ((size_t)lo_blk << LB);
  const unsigned lo_mask = (1u << LB) - 1;
  uint32_t* colp = out + col * n;

  for (unsigned e = threadIdx.x; e < tile; e += blockDim.x) {
    size_t g = base + ((size_t)(e >> LB) << s_lo) + (e & lo_mask);
    sm[e] = ebits ? in[col * (n >> ebits) + (g >> ebits)] : in[col * n + g];
  }
  __syncthreads();

  const unsigned pairs = tile >> 1;
  if (!INVERSE) {
    for (unsigned t = t_begin; t <= B; t++) {
      const unsigned half = 1u << (t - 1);
      const unsigned s = s_lo + t;
      for (unsigned p = threadIdx.x; p < pairs; p += blockDim.x) {
        unsigned lo = p & lo_mask, pm = p >> LB;
        unsigned jm = pm & (half - 1);
        unsigned m0 = ((pm >> (t - 1)) << t) | jm;
        unsigned i0 = (m0 << LB) | lo, i1 = i0 + (half << LB);
        size_t j = ((size_t)jm << s_lo) | ((size_t)lo_blk << LB) | lo;
        uint32_t w = tw[j << (logn - s)];
        uint32_t a = sm[i0], b = fp_mul(sm[i1], w);
        sm[i0] = fp_add(a, b);
        sm[i1] = fp_sub(a, b);
      }
      __syncthreads();
    }
  } else {
    for (unsigned t = B; t >= t_begin; t--) {
      const unsigned half = 1u << (t - 1);
      const unsigned s = s_lo + t;
      for (unsigned p = threadIdx.x; p < pairs; p += blockDim.x) {
        unsigned lo = p & lo_mask, pm = p >> LB;
        unsigned jm = pm & (half - 1);
        unsigned m0 = ((pm >> (t - 1)) << t) | jm;
        unsigned i0 = (m0 << LB) | lo, i1 = i0 + (half << LB);
        size_t j = ((size_t)jm << s_lo) | ((size_t)lo_blk << LB) | lo;
        uint32_t w = tw[j << (logn - s)];
        uint32_t a = sm[i0], b = sm[i1];
        sm[i0] = fp_add(a, b);
        sm[i1] = fp_mul(fp_sub(a, b), w);
      }
      __syncthreads();
    }
  }

  for (unsigned e = threadIdx.x; e < tile; e += blockDim.x) {
    size_t g = base + ((size_t)(e >> LB) << s_lo) + (e & lo_mask);
    uint32_t v = sm[e];
    if (apply_scale) {
      v = fp_mul(v, scale);
      if (scale_tab) v = fp_mul(v, scale_tab[g]);
    }
    colp[g] = v;
  }
}

struct NttPlan {
  unsigned T;  // log2 tile elements of the contiguous pass
  unsigned B2; // layers handled by the strided pass (0 = single pass)
  unsigned LB; // log2 contiguous run of the strided pass
  unsigned threads;
};
static NttPlan make_plan(unsigned logn) {
  NttPlan p;
  if (logn <= 13) {
    p.T = logn;
    p.B2 = 0;
    p.LB = 0;
  } else {
    unsigned need = (logn + 4 + 1) / 2;
    p.T = need < 13 ? 13 : need;
    if (p.T > 15) p.T = 15;
    p.B2 = logn - p.T;
    p.LB = p.T - p.B2;
  }
  p.threads = p.T <= 13 ? 256 : (p.T == 14 ? 512 : 1024);
  return p;
}

template <bool INV>
static b2z_err launch_pass(b2z_ctx* ctx, uint32_t* out, const uint32_t* in, const uint32_t* tw, const uint32_t* scale_tab,
                           uint32_t scale, unsigned logn, unsigned s_lo, unsigned B, unsigned LB, unsigned t_begin, unsigned ebits,
                           int apply_scale, size_t count, unsigned threads) {
  size_t tiles = ((size_t)1 << logn) >> (B + LB);
  size_t smem = sizeof(uint32_t) << (B + LB);
  static bool attr_set[2] = {false, false};
  if (smem > 48 * 1024 && !attr_set[INV]) {
    B2Z_CUDA(cudaFuncSetAttribute(ntt_pass_kernel<INV>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    attr_set[INV] = true;
  }
  // gridDim.y is limited to 65535 columns; the prover never exceeds a few thousand
  B2Z_REQUIRE(count <= 65535, "too many columns for one NTT launch");
  dim3 grid((unsigned)tiles, (unsigned)count);
  unsigned tile = 1u << (B + LB);
  unsigned th = threads;
  while (th > 32 && th > tile / 2) th >>= 1;
  ntt_pass_kernel<INV><<<grid, th, smem, ctx->stream>>>(out, in, tw, scale_tab, scale, logn, s_lo, B, LB, t_begin, ebits, apply_scale);
  ctx->launches++;
  B2Z_CUDA(cudaGetLastError());
  return nullptr;
}

b2z_err launch_interpolate(b2z_ctx* ctx, uint32_t* io, size_t count, unsigned po2, bool fuse_zk_shift) {
  if (count == 0) return nullptr;
  B2Z_TRY(ensure_tables(ctx, po2, false, true, fuse_zk_shift));
  uint32_t norm = fp_inv(fp_from_u32((uint32_t)((uint64_t)1 << po2) % P));
  const uint32_t* tab = fuse_zk_shift ? ctx->pow3rev[po2] : nullptr;
  if (po2 == 0) return nullptr;
  NttPlan p = make_plan(po2);
  if (p.B2) B2Z_TRY(launch_pass<true>(ctx, io, io, ctx->tw_rev[po2], nullptr, 0, po2, p.T, p.B2, p.LB, 1, 0, 0, count, p.threads));
  B2Z_TRY(launch_pass<true>(ctx, io, io, ctx->tw_rev[po2], tab, norm, po2, 0, p.T, 0, 1, 0, 1, count, p.threads));
  return nullptr;
}

b2z_err launch_evaluate(b2z_ctx* ctx, uint32_t* out, const uint32_t* in, size_t count, unsigned out_po2, unsigned expand_bits) {
  if (count == 0) return nullptr;
  B2Z_TRY(ensure_tables(ctx, out_po2, true, false, false));
  NttPlan p = make_plan(out_po2);
  B2Z_REQUIRE(expand_bits <= p.T, "expand_bits larger than the first tile");
  bool expanding = (in != out);
  // contiguous pass: layers expand_bits+1 .. T, loading through the expand map when `in` is the short array
  B2Z_TRY(launch_pass<false>(ctx, out, in, ctx->tw_fwd[out_po2], nullptr, 0, out_po2, 0, p.T, 0, expand_bits + 1,
                             expanding ? expand_bits : 0, 0, count, p.threads));
  if (p.B2) B2Z_TRY(launch_pass<false>(ctx, out, out, ctx->tw_fwd[out_po2], nullptr, 0, out_po2, p.T, p.B2, p.LB, 1, 0, 0, count, p.threads));
  return nullptr;
}

__global__ void bit_reverse_kernel(uint32_t* io, unsigned bits, size_t total) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= total) return;
  size_t col = i >> bits;
  uint32_t idx = (uint32_t)(i & (((size_t)1 << bits) - 1));
  uint32_t r = bit_rev(idx, bits);
  if (idx < r) {
    uint32_t* p = io + (col << bits);
    uint32_t a = p[idx], b = p[r];
    p[idx] = b;
    p[r] = a;
  }
}
b2z_err launch_bit_reverse(b2z_ctx* ctx, uint32_t* io, size_t count, unsigned po2) {
  size_t total = count << po2;
  if (!total || po2 == 0) return nullptr;
  bit_reverse_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(io, po2, total);
  ctx->launches++;
  B2Z_CUDA(cudaGetLastError());
  return nullptr;
}

__global__ void zk_shift_kernel(uint32_t* io, const uint32_t* __restrict__ tab, unsigned bits, size_t total) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= total) return;
  io[i] = fp_mul(io[i], tab[i & (((size_t)1 << bits) - 1)]);
}
b2z_err launch_zk_shift(b2z_ctx* ctx, uint32_t* io, size_t count, unsigned po2) {
  size_t total = count << po2;
  if (!total) return nullptr;
  B2Z_TRY(ensure_tables(ctx, po2, false, false, true));
  zk_shift_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(io, ctx->pow3rev[po2], po2, total);
  ctx->launches++;
  B2Z_CUDA(cudaGetLastError());
  return nullptr;
}

}  // namespace b2z
