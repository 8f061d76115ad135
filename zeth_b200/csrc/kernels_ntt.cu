// BabyBear NTT kernels: batch_interpolate_ntt (natural -> bit-reversed, 1/n, optional fused zk_shift),
// batch_expand_into_evaluate_ntt / batch_evaluate_ntt (bit-reversed -> natural), bit_reverse, zk_shift.
//
// Semantics follow risc0-zkp 3.0.2 core/ntt.rs + hal/cpu.rs (scope rows a5-a8); layout is column-major
// [count][n].  The implementation is NOT upstream's layer-by-layer recursion.  A 2^L-point transform is factored
// (Cooley-Tukey, L = R1 + R2 + ...) into passes of R <= 8 bits; pass k works on the bits above the previous passes and
// is preceded by one element-wise "twist" multiplication w^(rev(i_hi) * k_lo) (table lookup, coalesced), after which
// every 2^R-point sub-transform uses the SAME twiddles.  That makes all twiddles warp-uniform:
//   * a CTA stages a tile of [2^R points] x [32 lanes] in shared memory; lane = one independent sub-transform
//     (32 adjacent low-index positions for strided passes, 32 consecutive contiguous rows for the first pass, with a
//     padded pitch so that every shared-memory access of a warp is conflict-free);
//   * each thread runs radix-16 butterflies on 16 registers with the sixteenth roots of unity as constant-bank
//     operands; one uniform twist between the two radix-16 groups of a pass;
//   * the expand step (out[i] = in[i >> 2]) is folded into the first pass's loads, 1/n and the zk_shift factor
//     3^rev(j) into the last inverse pass's stores.
// Each pass is one launch over all columns (L2-sized column chunks were measured slower: the passes are instruction
// bound, not bandwidth bound).
#include <cuda.h>  // CUtensorMap types only; the encoder is fetched with cudaGetDriverEntryPoint (no libcuda link)

#include <map>
#include <mutex>
#include <tuple>

#include "common.h"

namespace b2z {

// sixteenth roots of unity (Montgomery form): c_w16[j] = w16^j, c_w16i[j] = w16^-j, j < 8
__constant__ uint32_t c_w16[8];
__constant__ uint32_t c_w16i[8];
__constant__ uint32_t c_w16p[8], c_w16s[8];    // plain (non-Montgomery) w16^j and its Shoup companion floor(w16^j * 2^32 / P)
__constant__ uint32_t c_w16ip[8], c_w16is[8];  // the same for w16^-j
static bool g_w16_uploaded[16] = {};

__global__ void build_pow3rev_kernel(uint32_t* tab, unsigned bits) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < ((size_t)1 << bits)) tab[i] = fp_pow(MONT_THREE, bit_rev((uint32_t)i, bits));
}
// twist table for a pass boundary: T[i] = w^(rev_b(i >> a) * (i & (2^a - 1))), w = primitive 2^(a+b)-th root (or inverse)
__global__ void build_twist_kernel(uint32_t* tab, uint32_t w, unsigned a, unsigned b) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i >= ((size_t)1 << (a + b))) return;
  uint64_t e = (uint64_t)bit_rev((uint32_t)(i >> a), b) * (uint64_t)(i & (((size_t)1 << a) - 1));
  tab[i] = fp_pow(w, e);
}

struct TwistKey {
  unsigned a, b;
  bool inv;
  bool operator<(const TwistKey& o) const { return std::tie(a, b, inv) < std::tie(o.a, o.b, o.inv); }
};
struct NttTables {
  std::map<TwistKey, uint32_t*> twist;
};
static std::map<b2z_ctx*, NttTables> g_tables;  // contexts are single-threaded, the registry itself is shared
static std::mutex g_tables_mutex;

static b2z_err get_twist(b2z_ctx* ctx, unsigned a, unsigned b, bool inv, const uint32_t** out) {
  std::lock_guard<std::mutex> lock(g_tables_mutex);
  NttTables& t = g_tables[ctx];
  TwistKey k{a, b, inv};
  auto it = t.twist.find(k);
  if (it == t.twist.end()) {
    uint32_t* p = nullptr;
    size_t n = (size_t)1 << (a + b);
    B2Z_CUDA(cudaMalloc(&p, 4 * n));
    uint32_t w = inv ? rou_rev(a + b) : rou_fwd(a + b);
    build_twist_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(p, w, a, b);
    ctx->launches++;
    B2Z_CUDA(cudaGetLastError());
    it = t.twist.emplace(k, p).first;
  }
  *out = it->second;
  return nullptr;
}
void ntt_release_tables(b2z_ctx* ctx) {
  std::lock_guard<std::mutex> lock(g_tables_mutex);
  auto it = g_tables.find(ctx);
  if (it == g_tables.end()) return;
  for (auto& kv : it->second.twist) cudaFree(kv.second);
  g_tables.erase(it);
}

b2z_err ensure_tables(b2z_ctx* ctx, unsigned po2, bool pow3) {
  B2Z_REQUIRE(po2 <= MAX_PO2, "transform size exceeds 2^26");
  static std::mutex w16_mutex;
  std::lock_guard<std::mutex> w16_lock(w16_mutex);
  if (ctx->device < 16 && !g_w16_uploaded[ctx->device]) {
    uint32_t w[8], wi[8], g = rou_fwd(4), gi = rou_rev(4);
    w[0] = wi[0] = MONT_ONE;
    for (int j = 1; j < 8; j++) {
      w[j] = fp_mul(w[j - 1], g);
      wi[j] = fp_mul(wi[j - 1], gi);
    }
    B2Z_CUDA(cudaMemcpyToSymbol(c_w16, w, sizeof w));
    B2Z_CUDA(cudaMemcpyToSymbol(c_w16i, wi, sizeof wi));
    uint32_t wp[8], ws[8], wip[8], wis[8];
    for (int j = 0; j < 8; j++) {
      wp[j] = fp_to_u32(w[j]);
      wip[j] = fp_to_u32(wi[j]);
      ws[j] = (uint32_t)((((unsigned __int128)wp[j]) << 32) / P);
      wis[j] = (uint32_t)((((unsigned __int128)wip[j]) << 32) / P);
    }
    B2Z_CUDA(cudaMemcpyToSymbol(c_w16p, wp, sizeof wp));
    B2Z_CUDA(cudaMemcpyToSymbol(c_w16s, ws, sizeof ws));
    B2Z_CUDA(cudaMemcpyToSymbol(c_w16ip, wip, sizeof wip));
    B2Z_CUDA(cudaMemcpyToSymbol(c_w16is, wis, sizeof wis));
    g_w16_uploaded[ctx->device] = true;
  }
  if (pow3 && !ctx->pow3rev[po2]) {
    size_t n = (size_t)1 << po2;
    B2Z_CUDA(cudaMalloc(&ctx->pow3rev[po2], sizeof(uint32_t) * n));
    build_pow3rev_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->pow3rev[po2], po2);
    ctx->launches++;
    B2Z_CUDA(cudaGetLastError());
  }
  return nullptr;
}


// ---- TMA (cp.async.bulk.tensor) staging of strided tiles ----
// A strided pass tile is a box {32 adjacent positions, 2^R points at stride 2^s_lo, 1 block} of the 3-D view
// [total / 2^(s_lo+R)][2^R][2^s_lo] of the column buffer; one elected thread issues the bulk tensor copy, the tile lands in
// shared memory as dense [point][lane] (exactly the compute layout) and completion is signalled on an mbarrier.
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
               "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"((uint64_t)map), "r"(smem_u32(src)), "r"(c0),
               "r"(c1), "r"(c2)
               : "memory");
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

// ---- register-level radix-2^Q butterflies with constant sixteenth-root twiddles ----
// Instruction diet (these kernels are issue-bound: ~200 instructions per element of a 2^24-point transform):
//   * constant twiddles are Shoup products (IMAD.HI + 2 IMAD + 1 correction = 4 instructions instead of the 5 of a
//     Montgomery product); a Shoup product accepts ANY 32-bit input, so
//   * a sum / difference whose only consumer is such a product (or the twist product that follows the group when LAZY_OUT)
//     is left unreduced in [0, 2P): a + b, resp. a - b + P -- one instruction instead of two.
// Every value that leaves a kernel is canonical; intermediate tiles may hold [0, 2P) words.
__device__ __forceinline__ uint32_t mul_shoup(uint32_t x, uint32_t w, uint32_t ws) {  // x < 2^32, w plain < P -> canonical x*w mod P
  uint32_t q = __umulhi(x, ws);
  uint32_t r = x * w - q * P;  // in [0, 2P)
  return umin32(r, r - P);
}
// forward: decimation in time, layers FIRST..Q (in place; positions hold the bit-reversed-order input)
template <int Q, int FIRST, bool LAZY_OUT = false>
__device__ __forceinline__ void dit_base(uint32_t* v) {
#pragma unroll
  for (int s = FIRST; s <= Q; s++) {
    const int half = 1 << (s - 1);
#pragma unroll
    for (int blk = 0; blk < (1 << Q); blk += 2 * half) {
#pragma unroll
      for (int j = 0; j < half; j++) {
        const int lo = blk + j, hi = blk + j + half;
        uint32_t a = v[lo];
        uint32_t b = v[hi];
        if (j != 0) b = mul_shoup(b, c_w16p[j << (4 - s)], c_w16s[j << (4 - s)]);
        // position p of the next layer is a multiplied `b` input iff it lies in the upper half of its block, off the block's first column
        const bool lazy_lo = s == Q ? LAZY_OUT : ((lo & (4 * half - 1)) >= 2 * half && (lo & (2 * half - 1)) != 0);
        const bool lazy_hi = s == Q ? LAZY_OUT : ((hi & (4 * half - 1)) >= 2 * half && (hi & (2 * half - 1)) != 0);
        v[lo] = lazy_lo ? a + b : fp_add(a, b);
        v[hi] = lazy_hi ? a - b + P : fp_sub(a, b);
      }
    }
  }
}
// inverse: decimation in frequency, layers Q..1, inverse twiddles (natural in -> bit-reversed out)
template <int Q, bool LAZY_OUT = false>
__device__ __forceinline__ void dif_base(uint32_t* v) {
#pragma unroll
  for (int s = Q; s >= 1; s--) {
    const int half = 1 << (s - 1);
#pragma unroll
    for (int blk = 0; blk < (1 << Q); blk += 2 * half) {
#pragma unroll
      for (int j = 0; j < half; j++) {
        uint32_t a = v[blk + j];
        uint32_t b = v[blk + j + half];
        const bool lazy = s == 1 && LAZY_OUT;
        v[blk + j] = lazy ? a + b : fp_add(a, b);
        if (j != 0) v[blk + j + half] = mul_shoup(a - b + P, c_w16ip[j << (4 - s)], c_w16is[j << (4 - s)]);
        else v[blk + j + half] = lazy ? a - b + P : fp_sub(a, b);
      }
    }
  }
}

// One radix group over bits [G_LO, G_LO + Q) of the R-bit point index, for all lanes of the tile.
//   smem address of (point p, lane l) = p * PITCH_P + l * PITCH_L  (all compile-time, so every access is base + immediate)
//   intra-pass twist (warp-uniform, staged in shared memory): multiply point p by tws[p] after (forward) / before (inverse)
// PTW: additionally apply the inter-pass twist here (TMA path: the tile arrives untouched in shared memory, so the twist
//      moves from the tile load/store loops into the first-executed / last-executed radix group); `ptw` points at the
//      table entry of (point 0, this lane), consecutive points are `pstride` entries apart, `pmask` wraps the period.
template <int R, int Q, int G_LO, bool ROWS, bool INV, int FIRST, bool TWIST, bool PTW = false>
__device__ __forceinline__ void radix_group(uint32_t* sm, const uint32_t* tws, unsigned lane, unsigned warp, unsigned nwarps,
                                            const uint32_t* __restrict__ ptab = nullptr, uint32_t pbase = 0, uint32_t pstride = 0,
                                            uint32_t pmask = 0) {
  constexpr unsigned PITCH_P = ROWS ? 1u : 32u;
  constexpr unsigned PITCH_L = ROWS ? ((1u << R) + 1u) : 1u;
  constexpr unsigned SETS = 1u << (R - Q);
  uint32_t* lane_base = sm + lane * PITCH_L;
  (void)nwarps;  // kernels always run 8 warps: the set loop has a compile-time trip count (no select/move code, no stack v[])
  constexpr unsigned ITER = SETS >= 8 ? SETS / 8 : 1;
#pragma unroll
  for (unsigned it = 0; it < ITER; it++) {
    const unsigned set = warp + 8 * it;
    if (SETS < 8 && set >= SETS) break;
    // set index -> the point bits outside [G_LO, G_LO+Q)
    const unsigned low = set & ((1u << G_LO) - 1), high = set >> G_LO;
    const unsigned p0 = (high << (G_LO + Q)) | low;
    uint32_t* base = lane_base + p0 * PITCH_P;
    uint32_t v[1 << Q];
    uint32_t pt[PTW ? (1 << Q) : 1];
    if (PTW) {
#pragma unroll
      for (int j = 0; j < (1 << Q); j++) pt[j] = __ldg(ptab + ((pbase + (p0 + (j << G_LO)) * pstride) & pmask));
    }
#pragma unroll
    for (int j = 0; j < (1 << Q); j++) {
      uint32_t x = base[(j << G_LO) * PITCH_P];
      if (PTW && !INV) x = fp_mul(x, pt[j]);
      if (INV && TWIST) x = fp_mul(x, tws[p0 + (j << G_LO)]);
      v[j] = x;
    }
    if (INV) dif_base<Q>(v);
    else dit_base<Q, FIRST>(v);
#pragma unroll
    for (int j = 0; j < (1 << Q); j++) {
      uint32_t x = v[j];
      if (!INV && TWIST) x = fp_mul(x, tws[p0 + (j << G_LO)]);
      if (PTW && INV) x = fp_mul(x, pt[j]);
      base[(j << G_LO) * PITCH_P] = x;
    }
  }
}

struct PassArgs {
  uint32_t* out;
  const uint32_t* in;
  const uint32_t* pass_twist;   // inter-pass twist table of 2^(s_lo + R) entries (nullptr for the first pass)
  const uint32_t* group_twist;  // intra-pass twist table of 2^R entries (nullptr when the pass has one group)
  const uint32_t* scale_tab;    // optional per-position factor on the final inverse store (zk_shift), period 2^logn
  uint32_t scale;               // final inverse scale (1/n), applied when apply_scale
  unsigned logn;                // per-column transform size
  unsigned s_lo;                // bits below this pass
  unsigned ebits;               // expand bits folded into the first forward pass's loads
  int apply_scale;
  size_t total;                 // total elements of the launch (columns * n)
};

// ROWS (s_lo == 0): tile = 32 consecutive rows of 2^R contiguous points, lane = row, pitch (2^R + 1)
// strided         : tile = 2^R points at stride 2^s_lo x 32 adjacent positions, lane = position, pitch_p = 32
template <bool INV, int FIRST, int R, bool ROWS>
__global__ void __launch_bounds__(256) ntt_pass_kernel(PassArgs a) {
  constexpr int Q1 = R < 4 ? R : 4;
  constexpr int Q2 = R - Q1;
  constexpr unsigned NPTS = 1u << R;
  constexpr unsigned PITCH_L = ROWS ? (NPTS + 1u) : 1u;
  constexpr unsigned TILE = NPTS * 32u;
  constexpr unsigned U = 8;
  extern __shared__ uint32_t sm[];
  uint32_t* tws = sm + (ROWS ? 32u * (NPTS + 1u) : TILE);  // intra-pass twist table (NPTS words)
  const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  size_t base;  // global index of (point 0, lane 0)
  if (ROWS) {
    base = (size_t)blockIdx.x * TILE;
  } else {
    const unsigned lo_blocks_log = a.s_lo - 5;
    size_t hi = (size_t)blockIdx.x >> lo_blocks_log;
    size_t lo_blk = (size_t)blockIdx.x & (((size_t)1 << lo_blocks_log) - 1);
    base = (hi << (a.s_lo + R)) | (lo_blk << 5);
  }
  const uint32_t twist_mask32 = (uint32_t)((((size_t)1 << (a.s_lo + R)) - 1));
  if (Q2 > 0) {
    for (unsigned p = threadIdx.x; p < NPTS; p += blockDim.x) tws[p] = a.group_twist[p];
  }

  // ---- load (forward: inter-pass twist applied here); U independent loads in flight per thread ----
  // element e = tid + 256*k of the tile: strided mode point p = (tid >> 5) + 8k, lane l = tid & 31;
  // rows mode row = (tid >> R) + k * 2^(8-R), column = tid & (NPTS-1)   (R <= 8)
  constexpr unsigned K = TILE >= 256 ? TILE / 256 : 1;   // elements per thread
  constexpr unsigned UU = K < U ? K : U;
  constexpr bool FULL = (TILE >= 256);                  // every thread has exactly K elements
  const unsigned tid = threadIdx.x;
  if (ROWS) {
    const unsigned sidx0 = (tid >> R) * PITCH_L + (tid & (NPTS - 1));
    constexpr unsigned SSTEP = (R <= 8 ? (256u >> R) : 0u) * PITCH_L;
#pragma unroll 1
    for (unsigned k0 = 0; k0 < K; k0 += UU) {
      uint32_t x[UU];
#pragma unroll
      for (unsigned u = 0; u < UU; u++) {
        unsigned e = tid + (k0 + u) * 256;
        size_t g = base + e;
        x[u] = 0;
        if ((FULL || e < TILE) && g < a.total) x[u] = a.ebits ? __ldg(a.in + (g >> a.ebits)) : a.in[g];
      }
#pragma unroll
      for (unsigned u = 0; u < UU; u++) {
        unsigned e = tid + (k0 + u) * 256;
        if (FULL || e < TILE) sm[sidx0 + (k0 + u) * SSTEP] = x[u];
      }
    }
  } else {
    const uint32_t* src = a.in + base;
    const uint32_t tbase = (uint32_t)base & twist_mask32;
    const uint32_t off0 = ((tid >> 5) << a.s_lo) + (tid & 31);
    const uint32_t ostep = 8u << a.s_lo;
    const bool tw_on = !INV && a.pass_twist;
#pragma unroll 1
    for (unsigned k0 = 0; k0 < K; k0 += UU) {
      uint32_t x[UU], tw[UU];
#pragma unroll
      for (unsigned u = 0; u < UU; u++) {
        unsigned e = tid + (k0 + u) * 256;
        x[u] = 0;
        tw[u] = 0;
        if (FULL || e < TILE) {
          uint32_t off = off0 + (k0 + u) * ostep;
          x[u] = src[off];
          if (tw_on) tw[u] = __ldg(a.pass_twist + ((tbase + off) & twist_mask32));
        }
      }
#pragma unroll
      for (unsigned u = 0; u < UU; u++) {
        unsigned e = tid + (k0 + u) * 256;
        if (FULL || e < TILE) sm[e] = tw_on ? fp_mul(x[u], tw[u]) : x[u];
      }
    }
  }
  __syncthreads();

  // ---- compute: at most two radix groups (R <= 8) ----
  if (!INV) {
    radix_group<R, Q1, 0, ROWS, false, FIRST, (Q2 > 0)>(sm, tws, lane, warp, nwarps);
    if (Q2 > 0) {
      __syncthreads();
      radix_group<R, (Q2 > 0 ? Q2 : 1), Q1, ROWS, false, 1, false>(sm, tws, lane, warp, nwarps);
    }
  } else {
    if (Q2 > 0) {
      radix_group<R, (Q2 > 0 ? Q2 : 1), Q1, ROWS, true, 1, false>(sm, tws, lane, warp, nwarps);
      __syncthreads();
    }
    radix_group<R, Q1, 0, ROWS, true, 1, (Q2 > 0)>(sm, tws, lane, warp, nwarps);
  }
  __syncthreads();

  // ---- store (inverse: inter-pass twist / final scale applied here) ----
  if (ROWS) {
    const uint32_t nmask = (uint32_t)(((size_t)1 << a.logn) - 1);
    const unsigned sidx0 = (tid >> R) * PITCH_L + (tid & (NPTS - 1));
    constexpr unsigned SSTEP = (R <= 8 ? (256u >> R) : 0u) * PITCH_L;
    const bool tab_on = a.apply_scale && a.scale_tab;
#pragma unroll 1
    for (unsigned k0 = 0; k0 < K; k0 += UU) {
      uint32_t tw[UU];
#pragma unroll
      for (unsigned u = 0; u < UU; u++) {
        unsigned e = tid + (k0 + u) * 256;
        size_t g = base + e;
        tw[u] = 0;
        if (tab_on && (FULL || e < TILE) && g < a.total) tw[u] = __ldg(a.scale_tab + ((uint32_t)g & nmask));
      }
#pragma unroll
      for (unsigned u = 0; u < UU; u++) {
        unsigned e = tid + (k0 + u) * 256;
        size_t g = base + e;
        if (!(FULL || e < TILE) || g >= a.total) continue;
        uint32_t x = sm[sidx0 + (k0 + u) * SSTEP];
        if (a.apply_scale) {
          x = fp_mul(x, a.scale);
          if (tab_on) x = fp_mul(x, tw[u]);
        }
        a.out[g] = x;
      }
    }
  } else {
    uint32_t* dst = a.out + base;
    const uint32_t tbase = (uint32_t)base & twist_mask32;
    const uint32_t off0 = ((tid >> 5) << a.s_lo) + (tid & 31);
    const uint32_t ostep = 8u << a.s_lo;
    const bool tw_on = INV && a.pass_twist;
#pragma unroll 1
    for (unsigned k0 = 0; k0 < K; k0 += UU) {
      uint32_t tw[UU];
#pragma unroll
      for (unsigned u = 0; u < UU; u++) {
        unsigned e = tid + (k0 + u) * 256;
        tw[u] = 0;
        if (tw_on && (FULL || e < TILE)) tw[u] = __ldg(a.pass_twist + ((tbase + off0 + (k0 + u) * ostep) & twist_mask32));
      }
#pragma unroll
      for (unsigned u = 0; u < UU; u++) {
        unsigned e = tid + (k0 + u) * 256;
        if (!(FULL || e < TILE)) continue;
        uint32_t x = sm[e];
        if (tw_on) x = fp_mul(x, tw[u]);
        dst[off0 + (k0 + u) * ostep] = x;
      }
    }
  }
}



// First forward pass of an x4 expand (out[i] = in[i >> 2], layers 1-2 skipped): the 16 points of a radix-16 set are
// 4 input words replicated 4 times, so each thread fetches them with ONE 128-bit load straight from global memory and
// the tile is never staged for the first radix group.
template <int R>
__global__ void __launch_bounds__(256) ntt_expand_first_kernel(PassArgs a) {
  constexpr int Q2 = R - 4;
  constexpr unsigned NPTS = 1u << R;
  constexpr unsigned PITCH_L = NPTS + 1u;
  constexpr unsigned TILE = NPTS * 32u;
  constexpr unsigned SETS = 1u << (R - 4);
  extern __shared__ uint32_t sm[];
  uint32_t* tws = sm + 32u * (NPTS + 1u);
  const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t base = (size_t)blockIdx.x * TILE;
  if (Q2 > 0) {
    for (unsigned p = tid; p < NPTS; p += blockDim.x) tws[p] = a.group_twist[p];
    __syncthreads();
  }
  const size_t row_start = base + (size_t)lane * NPTS;  // first output element of this lane's row
  const bool valid = row_start < a.total;
  const uint4* in4 = reinterpret_cast<const uint4*>(a.in) + (row_start >> 4);
  uint32_t* lane_base = sm + lane * PITCH_L;
  constexpr unsigned ITER = SETS >= 8 ? SETS / 8 : 1;
#pragma unroll
  for (unsigned it = 0; it < ITER; it++) {
    const unsigned set = warp + 8 * it;
    if (SETS < 8 && set >= SETS) break;
    uint4 q = valid ? __ldg(in4 + set) : make_uint4(0u, 0u, 0u, 0u);
    uint32_t v[16] = {q.x, q.x, q.x, q.x, q.y, q.y, q.y, q.y, q.z, q.z, q.z, q.z, q.w, q.w, q.w, q.w};
    dit_base<4, 3>(v);
    const unsigned p0 = set << 4;
#pragma unroll
    for (int j = 0; j < 16; j++) {
      uint32_t x = v[j];
      if (Q2 > 0) x = fp_mul(x, tws[p0 + j]);
      lane_base[p0 + j] = x;
    }
  }
  __syncthreads();
  if (Q2 > 0) {
    radix_group<R, (Q2 > 0 ? Q2 : 1), 4, true, false, 1, false>(sm, tws, lane, warp, 8);
    __syncthreads();
  }
  // store: element e = tid + 256 k -> row (tid >> R) + k * 2^(8-R), column tid & (NPTS - 1)
  constexpr unsigned K = TILE / 256;
  const unsigned sidx0 = (tid >> R) * PITCH_L + (tid & (NPTS - 1));
  constexpr unsigned SSTEP = (256u >> R) * PITCH_L;
  uint32_t* dst = a.out + base + tid;
  if (base + TILE <= a.total) {
#pragma unroll 8
    for (unsigned k = 0; k < K; k++) dst[(size_t)k * 256] = sm[sidx0 + k * SSTEP];
  } else {
    for (unsigned k = 0; k < K; k++)
      if (base + tid + (size_t)k * 256 < a.total) dst[(size_t)k * 256] = sm[sidx0 + k * SSTEP];
  }
}

// Strided pass with TMA staging (in place).  Same arithmetic as ntt_pass_kernel<INV, 1, R, false>.
template <bool INV, int R>
__global__ void __launch_bounds__(256) ntt_pass_tma_kernel(const __grid_constant__ CUtensorMap tmap, PassArgs a) {
  constexpr int Q1 = R < 4 ? R : 4;
  constexpr int Q2 = R - Q1;
  constexpr unsigned NPTS = 1u << R;
  constexpr unsigned TILE = NPTS * 32u;
  extern __shared__ __align__(128) uint32_t sm[];
  __shared__ __align__(8) uint64_t bar;
  uint32_t* tws = sm + TILE;
  const unsigned tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const unsigned lo_blocks_log = a.s_lo - 5;
  const unsigned hi = blockIdx.x >> lo_blocks_log;
  const unsigned lo_blk = blockIdx.x & ((1u << lo_blocks_log) - 1);
  if (tid == 0) mbar_init(&bar, 1);
  if (Q2 > 0) {
    for (unsigned p = tid; p < NPTS; p += blockDim.x) tws[p] = a.group_twist[p];
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&bar, TILE * 4);
    tma_load_3d(sm, &tmap, &bar, (int)(lo_blk << 5), 0, (int)hi);
  }
  // inter-pass twist table index of (point p, this lane) = (tbase + p * 2^s_lo) & mask  (the block index hi drops out)
  const uint32_t pmask = (uint32_t)((((size_t)1 << (a.s_lo + R)) - 1));
  const uint32_t tbase = (lo_blk << 5) + lane;
  const uint32_t pstride = 1u << a.s_lo;
  mbar_wait(&bar, 0);
  if (!INV) {
    radix_group<R, Q1, 0, false, false, 1, (Q2 > 0), true>(sm, tws, lane, warp, nwarps, a.pass_twist, tbase, pstride, pmask);
    if (Q2 > 0) {
      __syncthreads();
      radix_group<R, (Q2 > 0 ? Q2 : 1), Q1, false, false, 1, false>(sm, tws, lane, warp, nwarps);
    }
  } else {
    if (Q2 > 0) {
      radix_group<R, (Q2 > 0 ? Q2 : 1), Q1, false, true, 1, false>(sm, tws, lane, warp, nwarps);
      __syncthreads();
    }
    radix_group<R, Q1, 0, false, true, 1, (Q2 > 0), true>(sm, tws, lane, warp, nwarps, a.pass_twist, tbase, pstride, pmask);
  }
  // make the generic-proxy writes visible to the async proxy, then one thread stores the tile
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  __syncthreads();
  if (tid == 0) tma_store_3d(&tmap, sm, (int)(lo_blk << 5), 0, (int)hi);
}

// split logn into passes of <= 8 bits, first pass largest; every later pass needs s_lo >= 5 (32 adjacent lanes)
static void plan_passes(unsigned logn, unsigned* R, unsigned* np) {
  unsigned n = (logn + 7) / 8;
  if (n == 0) n = 1;
  if (n == 1) {
    R[0] = logn;
    *np = 1;
    return;
  }
  unsigned base = logn / n, rem = logn % n;
  for (unsigned i = 0; i < n; i++) R[i] = base + (i < rem ? 1 : 0);
  if (R[0] < 5) {  // logn = 9: 5 + 4
    unsigned need = 5 - R[0];
    R[0] += need;
    R[n - 1] -= need;
  }
  *np = n;
}

static size_t chunk_columns(size_t count, unsigned logn) {
  // Optional column chunking (B2Z_NTT_CHUNK_MB).  Keeping a chunk inside the 126 MB L2 was measured NOT to pay: the passes
  // are instruction bound, and one launch per pass over all columns is fastest (po2 20, 192 columns: 6.1 ms vs 7.2 ms
  // with 64 MB chunks, 9.7 ms with 16 MB chunks), so the default is a single chunk.
  size_t bytes_per_col = (size_t)4 << logn;
  static size_t chunk_mb = [] { const char* e = getenv("B2Z_NTT_CHUNK_MB"); return e ? (size_t)atol(e) : (size_t)(1u << 20); }();
  size_t c = (chunk_mb << 20) / bytes_per_col;
  if (c < 1) c = 1;
  return c < count ? c : count;
}

template <bool INV, int FIRST, int R, bool ROWS>
static b2z_err launch_inst(b2z_ctx* ctx, const PassArgs& a) {
  constexpr unsigned NPTS = 1u << R;
  size_t smem = sizeof(uint32_t) * ((ROWS ? 32 * (NPTS + 1) : 32 * NPTS) + NPTS);
  size_t tile_elems = (size_t)NPTS * 32;
  size_t tiles = (a.total + tile_elems - 1) / tile_elems;
  if (!ROWS) B2Z_REQUIRE(a.total % tile_elems == 0, "strided NTT pass needs whole tiles");
  ntt_pass_kernel<INV, FIRST, R, ROWS><<<(unsigned)tiles, 256, smem, ctx->stream>>>(a);
  ctx->launches++;
  B2Z_CUDA(cudaGetLastError());
  return nullptr;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encoder() {
  static EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
    return (EncodeTiledFn)p;
  }();
  return fn;
}
// B2Z_NTT_TMA=0 disables the TMA path (plain LDG/STG staging, same arithmetic)
static bool tma_enabled() {
  static bool on = [] {
    const char* e = getenv("B2Z_NTT_TMA");
    return !(e && e[0] == '0') && get_encoder() != nullptr;
  }();
  return on;
}
template <bool INV, int R>
static b2z_err launch_tma(b2z_ctx* ctx, const PassArgs& a) {
  constexpr unsigned NPTS = 1u << R;
  const size_t tile_elems = (size_t)NPTS * 32;
  B2Z_REQUIRE(a.total % tile_elems == 0 && a.in == a.out, "TMA NTT pass needs whole in-place tiles");
  CUtensorMap tmap;
  cuuint64_t dims[3] = {(cuuint64_t)1 << a.s_lo, NPTS, (cuuint64_t)(a.total >> (a.s_lo + R))};
  cuuint64_t strides[2] = {(cuuint64_t)4 << a.s_lo, (cuuint64_t)4 << (a.s_lo + R)};
  cuuint32_t box[3] = {32, NPTS, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult rc = get_encoder()(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, a.out, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  B2Z_REQUIRE(rc == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed");
  size_t smem = sizeof(uint32_t) * (32 * NPTS + NPTS);
  ntt_pass_tma_kernel<INV, R><<<(unsigned)(a.total / tile_elems), 256, smem, ctx->stream>>>(tmap, a);
  ctx->launches++;
  B2Z_CUDA(cudaGetLastError());
  return nullptr;
}

template <bool INV, int FIRST>
static b2z_err launch_one(b2z_ctx* ctx, const PassArgs& a, unsigned R) {
  const bool rows = (a.s_lo == 0);
  // first pass of an x4 expand: dedicated kernel (128-bit loads of the 4 distinct words of every radix-16 set)
  if (rows && !INV && FIRST == 3 && a.ebits == 2 && R >= 4 && R <= 8 && (reinterpret_cast<uintptr_t>(a.in) & 15) == 0) {
    size_t tile_elems = (size_t)32 << R;
    size_t tiles = (a.total + tile_elems - 1) / tile_elems;
    size_t smem = sizeof(uint32_t) * (32 * ((1u << R) + 1) + (1u << R));
    switch (R) {
      case 4: ntt_expand_first_kernel<4><<<(unsigned)tiles, 256, smem, ctx->stream>>>(a); break;
      case 5: ntt_expand_first_kernel<5><<<(unsigned)tiles, 256, smem, ctx->stream>>>(a); break;
      case 6: ntt_expand_first_kernel<6><<<(unsigned)tiles, 256, smem, ctx->stream>>>(a); break;
      case 7: ntt_expand_first_kernel<7><<<(unsigned)tiles, 256, smem, ctx->stream>>>(a); break;
      default: ntt_expand_first_kernel<8><<<(unsigned)tiles, 256, smem, ctx->stream>>>(a); break;
    }
    ctx->launches++;
    B2Z_CUDA(cudaGetLastError());
    return nullptr;
  }
  // strided in-place passes with >= 128-byte aligned rows go through TMA
  if (!rows && a.in == a.out && tma_enabled() && R >= 4 && (a.total >> (a.s_lo + R)) < (1ull << 31)) {
    switch (R) {
      case 4: return launch_tma<INV, 4>(ctx, a);
      case 5: return launch_tma<INV, 5>(ctx, a);
      case 6: return launch_tma<INV, 6>(ctx, a);
      case 7: return launch_tma<INV, 7>(ctx, a);
      case 8: return launch_tma<INV, 8>(ctx, a);
      default: break;
    }
  }
#define B2Z_NTT_CASE(RR)                                                         \
  case RR:                                                                       \
    return rows ? launch_inst<INV, FIRST, RR, true>(ctx, a) : launch_inst<INV, 1, RR, false>(ctx, a);
  switch (R) {
    B2Z_NTT_CASE(1) B2Z_NTT_CASE(2) B2Z_NTT_CASE(3) B2Z_NTT_CASE(4) B2Z_NTT_CASE(5) B2Z_NTT_CASE(6) B2Z_NTT_CASE(7) B2Z_NTT_CASE(8)
    default: return set_error("unsupported NTT pass width");
  }
#undef B2Z_NTT_CASE
}

}  // namespace b2z
#include "ntt_big.cuh"
namespace b2z {

// forward transform of `count` columns: out (2^logn per column) from in (2^(logn - ebits) per column when expanding)
b2z_err launch_evaluate(b2z_ctx* ctx, uint32_t* out, const uint32_t* in, size_t count, unsigned out_po2, unsigned expand_bits) {
  if (count == 0) return nullptr;
  B2Z_TRY(ensure_tables(ctx, out_po2, false));
  const unsigned logn = out_po2;
  bool expanding = (in != out);
  if (logn == 0) {
    if (expanding) B2Z_CUDA(cudaMemcpyAsync(out, in, 4 * count, cudaMemcpyDeviceToDevice, ctx->stream));
    return nullptr;
  }
  // 2^20 .. 2^24 points: two round trips through HBM instead of three (ntt_big.cuh)
  if (big_applies(logn) && ((expanding && expand_bits == 2 && (reinterpret_cast<uintptr_t>(in) & 15) == 0) || (!expanding && expand_bits == 0)))
    return launch_big_evaluate(ctx, out, in, count, logn, expanding);
  unsigned R[8], np = 0;
  plan_passes(logn, R, &np);
  B2Z_REQUIRE(expand_bits <= 2 && expand_bits <= (R[0] < 4 ? R[0] : 4), "expand_bits larger than the first radix group");
  const uint32_t* pass_tw[8] = {};
  const uint32_t* grp_tw[8] = {};
  unsigned s_lo = 0;
  for (unsigned k = 0; k < np; k++) {
    unsigned q1 = R[k] < 4 ? R[k] : 4;
    if (k > 0) B2Z_TRY(get_twist(ctx, s_lo, R[k], false, &pass_tw[k]));
    if (R[k] > q1) B2Z_TRY(get_twist(ctx, q1, R[k] - q1, false, &grp_tw[k]));
    s_lo += R[k];
  }
  const size_t n = (size_t)1 << logn;
  const size_t cc = chunk_columns(count, logn);
  for (size_t c0 = 0; c0 < count; c0 += cc) {
    size_t cols = (count - c0 < cc) ? count - c0 : cc;
    s_lo = 0;
    for (unsigned k = 0; k < np; k++) {
      PassArgs a{};
      a.out = out + c0 * n;
      a.in = (k == 0 && expanding) ? in + c0 * (n >> expand_bits) : out + c0 * n;
      a.pass_twist = pass_tw[k];
      a.group_twist = grp_tw[k];
      a.scale_tab = nullptr;
      a.scale = 0;
      a.logn = logn;
      a.s_lo = s_lo;
      a.ebits = (k == 0 && expanding) ? expand_bits : 0;
      a.apply_scale = 0;
      a.total = cols * n;
      unsigned first = (k == 0) ? expand_bits + 1 : 1;  // skipped layers are valid for in-place evaluate_ntt(expand_bits) too
      switch (first) {
        case 1: B2Z_TRY((launch_one<false, 1>(ctx, a, R[k]))); break;
        case 2: B2Z_TRY((launch_one<false, 2>(ctx, a, R[k]))); break;
        case 3: B2Z_TRY((launch_one<false, 3>(ctx, a, R[k]))); break;
        default: return set_error("expand_bits > 2 is not supported");
      }
      s_lo += R[k];
    }
  }
  return nullptr;
}

b2z_err launch_interpolate(b2z_ctx* ctx, uint32_t* io, size_t count, unsigned po2, bool fuse_zk_shift, const uint32_t* src) {
  if (count == 0) return nullptr;
  if (src == io) src = nullptr;
  B2Z_TRY(ensure_tables(ctx, po2, fuse_zk_shift));
  const unsigned logn = po2;
  if (logn == 0) {
    if (src) B2Z_CUDA(cudaMemcpyAsync(io, src, 4 * count, cudaMemcpyDeviceToDevice, ctx->stream));
    return nullptr;
  }
  uint32_t norm = fp_inv(fp_from_u32((uint32_t)(((uint64_t)1 << po2) % P)));
  if (big_applies(logn)) return launch_big_interpolate(ctx, io, src, count, logn, norm, fuse_zk_shift ? ctx->pow3rev[po2] : nullptr);
  unsigned R[8], np = 0;
  plan_passes(logn, R, &np);
  const uint32_t* pass_tw[8] = {};
  const uint32_t* grp_tw[8] = {};
  unsigned s_los[8];
  unsigned s_lo = 0;
  for (unsigned k = 0; k < np; k++) {
    unsigned q1 = R[k] < 4 ? R[k] : 4;
    s_los[k] = s_lo;
    if (k > 0) B2Z_TRY(get_twist(ctx, s_lo, R[k], true, &pass_tw[k]));
    if (R[k] > q1) B2Z_TRY(get_twist(ctx, q1, R[k] - q1, true, &grp_tw[k]));
    s_lo += R[k];
  }
  const size_t n = (size_t)1 << logn;
  const size_t cc = chunk_columns(count, logn);
  for (size_t c0 = 0; c0 < count; c0 += cc) {
    size_t cols = (count - c0 < cc) ? count - c0 : cc;
    for (unsigned kk = np; kk-- > 0;) {
      PassArgs a{};
      a.out = io + c0 * n;
      a.in = (src && kk == np - 1) ? src + c0 * n : io + c0 * n;
      a.pass_twist = pass_tw[kk];
      a.group_twist = grp_tw[kk];
      a.scale_tab = (kk == 0 && fuse_zk_shift) ? ctx->pow3rev[po2] : nullptr;
      a.scale = norm;
      a.logn = logn;
      a.s_lo = s_los[kk];
      a.ebits = 0;
      a.apply_scale = (kk == 0);
      a.total = cols * n;
      B2Z_TRY((launch_one<true, 1>(ctx, a, R[kk])));
    }
  }
  return nullptr;
}

// ---- bit reversal and stand-alone zk_shift ----
__global__ void bit_reverse_small_kernel(uint32_t* io, unsigned bits, size_t total) {
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
// bits >= 10: index = (hi:5)(mid)(lo:5).  The 32x32 tile {hi, lo} of a given mid maps onto the tile of rev(mid) with
// (hi', lo') = (rev lo, rev hi): a CTA swaps the two tiles through shared memory with coalesced 128-byte rows.
__global__ void __launch_bounds__(256) bit_reverse_tile_kernel(uint32_t* io, unsigned bits) {
  __shared__ uint32_t sa[32][33], sb[32][33];
  const unsigned mbits = bits - 10;
  const uint32_t mid = blockIdx.x, rmid = bit_rev(mid, mbits);
  if (mid > rmid) return;
  uint32_t* p = io + ((size_t)blockIdx.y << bits);
  const unsigned c = threadIdx.x & 31, r0 = threadIdx.x >> 5;
  for (unsigned r = r0; r < 32; r += 8) {
    sa[r][c] = p[((size_t)r << (bits - 5)) | ((size_t)mid << 5) | c];
    if (mid != rmid) sb[r][c] = p[((size_t)r << (bits - 5)) | ((size_t)rmid << 5) | c];
  }
  __syncthreads();
  const unsigned rc = bit_rev(c, 5);
  for (unsigned r = r0; r < 32; r += 8) {
    const unsigned rr = bit_rev(r, 5);
    if (mid != rmid) {
      p[((size_t)r << (bits - 5)) | ((size_t)rmid << 5) | c] = sa[rc][rr];
      p[((size_t)r << (bits - 5)) | ((size_t)mid << 5) | c] = sb[rc][rr];
    } else {
      p[((size_t)r << (bits - 5)) | ((size_t)mid << 5) | c] = sa[rc][rr];
    }
  }
}
b2z_err launch_bit_reverse(b2z_ctx* ctx, uint32_t* io, size_t count, unsigned po2) {
  size_t total = count << po2;
  if (!total || po2 == 0) return nullptr;
  if (po2 < 10 || count > 65535) {
    bit_reverse_small_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(io, po2, total);
  } else {
    bit_reverse_tile_kernel<<<dim3(1u << (po2 - 10), (unsigned)count), 256, 0, ctx->stream>>>(io, po2);
  }
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
  B2Z_TRY(ensure_tables(ctx, po2, true));
  zk_shift_kernel<<<(unsigned)((total + 255) / 256), 256, 0, ctx->stream>>>(io, ctx->pow3rev[po2], po2, total);
  ctx->launches++;
  B2Z_CUDA(cudaGetLastError());
  return nullptr;
}

}  // namespace b2z
