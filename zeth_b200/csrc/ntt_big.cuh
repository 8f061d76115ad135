// Two-pass BabyBear transforms for 2^20 .. 2^24 points (included by kernels_ntt.cu; same semantics as the multi-pass
// kernels there: risc0-zkp core/ntt.rs + hal batch_expand_into_evaluate_ntt / batch_interpolate_ntt).
//
// Why: with passes of <= 8 bits a 2^24-point LDE needs three round trips through HBM (84 B moved per input element
// against 20 B compulsory) and ~196 instructions per output element.  Here the transform is factored as
//     2^logn = 2^R1 (contiguous "row" pass, R1 = logn - 10 in 10..14)  x  2^10 (strided pass over 16 or 32 adjacent positions)
// so the data makes TWO round trips (52 B per input element for the LDE) and every element passes through fewer
// load / twist / store steps (~140 instructions per output element).
//   * A pass runs its radix-16 groups over a tile that stays in shared memory (row pass: one 2^R1-point row = up to 64 KB;
//     strided pass: 1024 points x LANES positions); the FIRST group of a pass takes its inputs straight from global memory
//     into registers and the LAST group stores straight from registers, so a tile crosses shared memory NG - 1 times.
//   * Shared-memory layout A(p) = p + (p >> 4) (one pad word per 16 points, in units of LANES words for the strided pass):
//     every group's accesses are `base + compile-time offset`, conflict-free for the groups over bits 0-3 and 4-7 and with
//     a single 2-way collision per warp access above that.
//   * Between groups every element is multiplied once by the Cooley-Tukey twist T(a, b)[p mod 2^(a+b)] (same tables as the
//     multi-pass path, get_twist); between the two passes by T(R1, 10)[index in column].  The strided pass enumerates its
//     CTAs column-fastest, so the 128 KB slice of the 2^logn-entry pass-twist table that a position block needs stays in
//     L2 for all columns (the multi-pass path streams that table once per column).
//   * The x4 expand (out[i] = in[i >> 2]) is folded into the row pass's first group (one 128-bit load feeds a radix-16
//     set, butterfly layers 1-2 skipped); 1/n and the zk_shift factor into the inverse row pass's 128-bit stores.
#pragma once

namespace b2z {

template <int R>
struct BigPlan {
  static constexpr int NG = (R + 3) / 4;          // radix groups of 4 bits, the last one of QL = R - 4 (NG - 1) bits
  static constexpr int QL = R - 4 * (NG - 1);
  static constexpr int ROWS = R >= 13 ? 1 : (1 << (13 - R));  // rows per CTA of the row pass
  static constexpr int E = ROWS << R;             // elements per CTA
  static constexpr int THREADS = E / 32;          // 32 elements (two radix-16 sets) per thread and group
  static constexpr int PITCH = (1 << R) + (1 << (R - 4));
  static constexpr int MINB = R >= 14 ? 2 : 4;
  static constexpr size_t SMEM = sizeof(uint32_t) * (size_t)ROWS * PITCH;
  __host__ __device__ static constexpr int q(int g) { return g == NG - 1 ? QL : 4; }
};

struct BigArgs {
  uint32_t* out;
  const uint32_t* in;
  const uint32_t* gt[3];       // twist between group g and g + 1 of this pass: T(4 (g + 1), q(g + 1)), indexed by p mod 2^(4 (g + 1) + q(g + 1))
  const uint32_t* pass_twist;  // strided pass: T(s_lo, R), indexed by the position inside the column
  const uint32_t* scale_tab;   // inverse row pass: optional per-position factor (zk_shift), period 2^logn
  uint32_t scale;              // inverse row pass: 1/n
  unsigned logn, s_lo;
  unsigned cols;
};

__device__ __forceinline__ void ld16(uint32_t* v, const uint32_t* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
#pragma unroll
  for (int k = 0; k < 4; k++) {
    uint4 x = __ldg(q + k);
    v[4 * k] = x.x, v[4 * k + 1] = x.y, v[4 * k + 2] = x.z, v[4 * k + 3] = x.w;
  }
}

// ---- row pass, forward (bit-reversed in -> natural out; expand folded into the loads) ----
template <int R, int FIRST, bool EXPAND>
__global__ void __launch_bounds__(BigPlan<R>::THREADS, BigPlan<R>::MINB) ntt_big_rows_fwd_kernel(BigArgs a) {
  using PL = BigPlan<R>;
  extern __shared__ uint32_t sm[];
  const unsigned tid = threadIdx.x;
  const size_t row0 = (size_t)blockIdx.x * PL::ROWS;
  constexpr unsigned SETS_ROW = 1u << (R - 4);
  // group 0 (bits 0-3): global -> registers -> shared
  {
    constexpr unsigned M = (1u << (4 + PL::q(1))) - 1;
#pragma unroll 1
    for (int it = 0; it < 2; it++) {
      const unsigned s = tid + it * PL::THREADS;
      const unsigned rl = s >> (R - 4), sr = s & (SETS_ROW - 1);
      const size_t grow = row0 + rl;
      uint32_t v[16], tw[16];
      if (EXPAND) {
        const uint4 q = __ldg(reinterpret_cast<const uint4*>(a.in) + grow * SETS_ROW + sr);
        v[0] = v[1] = v[2] = v[3] = q.x;
        v[4] = v[5] = v[6] = v[7] = q.y;
        v[8] = v[9] = v[10] = v[11] = q.z;
        v[12] = v[13] = v[14] = v[15] = q.w;
      } else {
        ld16(v, a.in + (grow << R) + 16 * sr);
      }
      ld16(tw, a.gt[0] + ((16 * sr) & M));
      dit_base<4, FIRST, true>(v);  // outputs in [0, 2P): the twist product below reduces them
      uint32_t* dst = sm + rl * PL::PITCH + 17 * sr;
#pragma unroll
      for (int j = 0; j < 16; j++) dst[j] = fp_mul(v[j], tw[j]);
    }
  }
  __syncthreads();
  // middle groups: shared -> shared
#pragma unroll
  for (int g = 1; g < PL::NG - 1; g++) {
    const int LO = 4 * g;
    const unsigned SJ = (1u << LO) + (1u << (LO - 4));
    const unsigned M = (1u << (LO + 4 + PL::q(g + 1))) - 1;
#pragma unroll 1
    for (int it = 0; it < 2; it++) {
      const unsigned s = tid + it * PL::THREADS;
      const unsigned rl = s >> (R - 4), sr = s & (SETS_ROW - 1);
      const unsigned low = sr & ((1u << LO) - 1), high = sr >> LO;
      const unsigned p0 = (high << (LO + 4)) | low;
      uint32_t* base = sm + rl * PL::PITCH + p0 + (p0 >> 4);
      const uint32_t* twp = a.gt[g] + (p0 & M);
      uint32_t v[16], tw[16];
#pragma unroll
      for (int j = 0; j < 16; j++) tw[j] = __ldg(twp + (j << LO));
#pragma unroll
      for (int j = 0; j < 16; j++) v[j] = base[j * SJ];
      dit_base<4, 1, true>(v);
#pragma unroll
      for (int j = 0; j < 16; j++) base[j * SJ] = fp_mul(v[j], tw[j]);
    }
    __syncthreads();
  }
  // last group: shared -> registers -> global (lanes = consecutive positions: coalesced)
  {
    constexpr int LO = 4 * (PL::NG - 1), Q = PL::QL;
    constexpr unsigned SJ = (1u << LO) + (1u << (LO - 4));
#pragma unroll 1
    for (int it = 0; it < (32 >> Q); it++) {
      const unsigned s = tid + it * PL::THREADS;
      const unsigned rl = s >> LO, low = s & ((1u << LO) - 1);
      const uint32_t* base = sm + rl * PL::PITCH + low + (low >> 4);
      uint32_t v[1 << Q];
#pragma unroll
      for (int j = 0; j < (1 << Q); j++) v[j] = base[j * SJ];
      dit_base<Q, 1, true>(v);  // [0, 2P) words between the passes: the strided pass multiplies by its pass twist first
      uint32_t* dst = a.out + ((row0 + rl) << R) + low;
#pragma unroll
      for (int j = 0; j < (1 << Q); j++) dst[(size_t)j << LO] = v[j];
    }
  }
}

// ---- row pass, inverse (natural in -> bit-reversed out; 1/n and the zk_shift table folded into the stores) ----
template <int R>
__global__ void __launch_bounds__(BigPlan<R>::THREADS, BigPlan<R>::MINB) ntt_big_rows_inv_kernel(BigArgs a) {
  using PL = BigPlan<R>;
  extern __shared__ uint32_t sm[];
  const unsigned tid = threadIdx.x;
  const size_t row0 = (size_t)blockIdx.x * PL::ROWS;
  constexpr unsigned SETS_ROW = 1u << (R - 4);
  // top group: global -> registers -> shared
  {
    constexpr int LO = 4 * (PL::NG - 1), Q = PL::QL;
    constexpr unsigned SJ = (1u << LO) + (1u << (LO - 4));
#pragma unroll 1
    for (int it = 0; it < (32 >> Q); it++) {
      const unsigned s = tid + it * PL::THREADS;
      const unsigned rl = s >> LO, low = s & ((1u << LO) - 1);
      const uint32_t* src = a.in + ((row0 + rl) << R) + low;
      uint32_t v[1 << Q];
#pragma unroll
      for (int j = 0; j < (1 << Q); j++) v[j] = src[(size_t)j << LO];
      dif_base<Q, true>(v);
      uint32_t* base = sm + rl * PL::PITCH + low + (low >> 4);
#pragma unroll
      for (int j = 0; j < (1 << Q); j++) base[j * SJ] = v[j];
    }
  }
  __syncthreads();
#pragma unroll
  for (int g = PL::NG - 2; g >= 1; g--) {
    const int LO = 4 * g;
    const unsigned SJ = (1u << LO) + (1u << (LO - 4));
    const unsigned M = (1u << (LO + 4 + PL::q(g + 1))) - 1;
#pragma unroll 1
    for (int it = 0; it < 2; it++) {
      const unsigned s = tid + it * PL::THREADS;
      const unsigned rl = s >> (R - 4), sr = s & (SETS_ROW - 1);
      const unsigned low = sr & ((1u << LO) - 1), high = sr >> LO;
      const unsigned p0 = (high << (LO + 4)) | low;
      uint32_t* base = sm + rl * PL::PITCH + p0 + (p0 >> 4);
      const uint32_t* twp = a.gt[g] + (p0 & M);
      uint32_t v[16], tw[16];
#pragma unroll
      for (int j = 0; j < 16; j++) tw[j] = __ldg(twp + (j << LO));
#pragma unroll
      for (int j = 0; j < 16; j++) v[j] = fp_mul(base[j * SJ], tw[j]);
      dif_base<4, true>(v);
#pragma unroll
      for (int j = 0; j < 16; j++) base[j * SJ] = v[j];
    }
    __syncthreads();
  }
  // group 0: shared -> registers -> global, 16 contiguous words per thread
  {
    constexpr unsigned M = (1u << (4 + PL::q(1))) - 1;
    const uint32_t nmask = (uint32_t)(((size_t)1 << a.logn) - 1);
#pragma unroll 1
    for (int it = 0; it < 2; it++) {
      const unsigned s = tid + it * PL::THREADS;
      const unsigned rl = s >> (R - 4), sr = s & (SETS_ROW - 1);
      const size_t g0 = ((row0 + rl) << R) + 16 * sr;
      uint32_t v[16], tw[16];
      ld16(tw, a.gt[0] + ((16 * sr) & M));
      const uint32_t* src = sm + rl * PL::PITCH + 17 * sr;
#pragma unroll
      for (int j = 0; j < 16; j++) v[j] = fp_mul(src[j], tw[j]);
      dif_base<4, true>(v);
      if (a.scale_tab) {
        ld16(tw, a.scale_tab + ((uint32_t)g0 & nmask));
#pragma unroll
        for (int j = 0; j < 16; j++) v[j] = fp_mul(fp_mul(v[j], a.scale), tw[j]);
      } else {
#pragma unroll
        for (int j = 0; j < 16; j++) v[j] = fp_mul(v[j], a.scale);
      }
      uint4* dst = reinterpret_cast<uint4*>(a.out + g0);
#pragma unroll
      for (int k = 0; k < 4; k++) dst[k] = make_uint4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
    }
  }
}

// ---- strided pass: 2^R points at stride 2^s_lo x LANES adjacent positions per CTA ----
template <bool INV, int R, int LANES>
__global__ void __launch_bounds__((LANES << R) / 32, LANES == 16 ? 2 : 1) ntt_big_strided_kernel(BigArgs a) {
  using PL = BigPlan<R>;
  constexpr int THREADS = (LANES << R) / 32;
  constexpr int TL = THREADS / LANES;  // threads per lane
  extern __shared__ uint32_t sm[];
  const unsigned tid = threadIdx.x, l = tid % LANES, ts = tid / LANES;
  const unsigned col = blockIdx.x % a.cols, kblk = blockIdx.x / a.cols;
  const unsigned k = kblk * LANES + l;
  const size_t colbase = (size_t)col << a.logn;
  const uint32_t* in = a.in + colbase + k;
  uint32_t* out = a.out + colbase + k;
  const uint32_t* pt = a.pass_twist + k;
  const unsigned s_lo = a.s_lo;
  if (!INV) {
    {  // group 0: global (+ pass twist) -> shared
      constexpr unsigned M = (1u << (4 + PL::q(1))) - 1;
#pragma unroll 1
      for (int it = 0; it < 2; it++) {
        const unsigned sl = ts + it * TL;
        uint32_t v[16], tw[16];
#pragma unroll
        for (int j = 0; j < 16; j++) {
          const size_t off = (size_t)(16 * sl + j) << s_lo;
          v[j] = in[off];
          tw[j] = __ldg(pt + off);
        }
#pragma unroll
        for (int j = 0; j < 16; j++) v[j] = fp_mul(v[j], tw[j]);
        dit_base<4, 1, true>(v);
        const uint32_t* g0 = a.gt[0] + ((16 * sl) & M);
        uint32_t* dst = sm + (17 * sl) * LANES + l;
#pragma unroll
        for (int j = 0; j < 16; j++) dst[j * LANES] = fp_mul(v[j], __ldg(g0 + j));
      }
    }
    __syncthreads();
#pragma unroll
    for (int g = 1; g < PL::NG - 1; g++) {
      const int LO = 4 * g;
      const unsigned SJ = ((1u << LO) + (1u << (LO - 4))) * LANES;
      const unsigned M = (1u << (LO + 4 + PL::q(g + 1))) - 1;
#pragma unroll 1
      for (int it = 0; it < 2; it++) {
        const unsigned sl = ts + it * TL;
        const unsigned low = sl & ((1u << LO) - 1), high = sl >> LO;
        const unsigned p0 = (high << (LO + 4)) | low;
        uint32_t* base = sm + (p0 + (p0 >> 4)) * LANES + l;
        const uint32_t* twp = a.gt[g] + (p0 & M);
        uint32_t v[16];
#pragma unroll
        for (int j = 0; j < 16; j++) v[j] = base[j * SJ];
        dit_base<4, 1, true>(v);
#pragma unroll
        for (int j = 0; j < 16; j++) base[j * SJ] = fp_mul(v[j], __ldg(twp + (j << LO)));
      }
      __syncthreads();
    }
    {  // last group: shared -> global
      constexpr int LO = 4 * (PL::NG - 1), Q = PL::QL;
      constexpr unsigned SJ = ((1u << LO) + (1u << (LO - 4))) * LANES;
#pragma unroll 1
      for (int it = 0; it < (32 >> Q); it++) {
        const unsigned low = ts + it * TL;
        const uint32_t* base = sm + (low + (low >> 4)) * LANES + l;
        uint32_t v[1 << Q];
#pragma unroll
        for (int j = 0; j < (1 << Q); j++) v[j] = base[j * SJ];
        dit_base<Q, 1>(v);
#pragma unroll
        for (int j = 0; j < (1 << Q); j++) out[(size_t)(low + (j << LO)) << s_lo] = v[j];
      }
    }
  } else {
    {  // top group: global -> shared
      constexpr int LO = 4 * (PL::NG - 1), Q = PL::QL;
      constexpr unsigned SJ = ((1u << LO) + (1u << (LO - 4))) * LANES;
#pragma unroll 1
      for (int it = 0; it < (32 >> Q); it++) {
        const unsigned low = ts + it * TL;
        uint32_t v[1 << Q];
#pragma unroll
        for (int j = 0; j < (1 << Q); j++) v[j] = in[(size_t)(low + (j << LO)) << s_lo];
        dif_base<Q, true>(v);
        uint32_t* base = sm + (low + (low >> 4)) * LANES + l;
#pragma unroll
        for (int j = 0; j < (1 << Q); j++) base[j * SJ] = v[j];
      }
    }
    __syncthreads();
#pragma unroll
    for (int g = PL::NG - 2; g >= 1; g--) {
      const int LO = 4 * g;
      const unsigned SJ = ((1u << LO) + (1u << (LO - 4))) * LANES;
      const unsigned M = (1u << (LO + 4 + PL::q(g + 1))) - 1;
#pragma unroll 1
      for (int it = 0; it < 2; it++) {
        const unsigned sl = ts + it * TL;
        const unsigned low = sl & ((1u << LO) - 1), high = sl >> LO;
        const unsigned p0 = (high << (LO + 4)) | low;
        uint32_t* base = sm + (p0 + (p0 >> 4)) * LANES + l;
        const uint32_t* twp = a.gt[g] + (p0 & M);
        uint32_t v[16];
#pragma unroll
        for (int j = 0; j < 16; j++) v[j] = fp_mul(base[j * SJ], __ldg(twp + (j << LO)));
        dif_base<4, true>(v);
#pragma unroll
        for (int j = 0; j < 16; j++) base[j * SJ] = v[j];
      }
      __syncthreads();
    }
    {  // group 0: shared -> (pass twist) -> global
      constexpr unsigned M = (1u << (4 + PL::q(1))) - 1;
#pragma unroll 1
      for (int it = 0; it < 2; it++) {
        const unsigned sl = ts + it * TL;
        const uint32_t* g0 = a.gt[0] + ((16 * sl) & M);
        const uint32_t* src = sm + (17 * sl) * LANES + l;
        uint32_t v[16], tw[16];
#pragma unroll
        for (int j = 0; j < 16; j++) tw[j] = __ldg(pt + ((size_t)(16 * sl + j) << s_lo));
#pragma unroll
        for (int j = 0; j < 16; j++) v[j] = fp_mul(src[j * LANES], __ldg(g0 + j));
        dif_base<4, true>(v);
#pragma unroll
        for (int j = 0; j < 16; j++) out[(size_t)(16 * sl + j) << s_lo] = fp_mul(v[j], tw[j]);
      }
    }
  }
}

// ---- host side ----
constexpr int BIG_R2 = 10;
static int big_lanes() {
  static int v = [] {
    const char* e = getenv("B2Z_NTT_BIG_LANES");
    int x = e ? atoi(e) : 16;
    return x == 32 ? 32 : 16;
  }();
  return v;
}
// B2Z_NTT_BIG=0 falls back to the multi-pass kernels for every size
static bool big_enabled() {
  static bool on = [] {
    const char* e = getenv("B2Z_NTT_BIG");
    return !(e && e[0] == '0');
  }();
  return on;
}
static bool big_applies(unsigned logn) { return big_enabled() && logn >= 20 && logn <= 24; }

template <int R>
static b2z_err big_tables(b2z_ctx* ctx, bool inv, const uint32_t** gt) {
  using PL = BigPlan<R>;
  for (int g = 0; g + 1 < PL::NG; g++) B2Z_TRY(get_twist(ctx, 4 * (g + 1), PL::q(g + 1), inv, &gt[g]));
  return nullptr;
}

template <bool INV>
static b2z_err launch_big_strided(b2z_ctx* ctx, BigArgs a) {
  B2Z_TRY(big_tables<BIG_R2>(ctx, INV, a.gt));
  B2Z_TRY(get_twist(ctx, a.s_lo, BIG_R2, INV, &a.pass_twist));
  const int lanes = big_lanes();
  const size_t positions = (size_t)1 << a.s_lo;
  const size_t grid = (positions / lanes) * a.cols;
  B2Z_REQUIRE(grid < (1ull << 31), "two-pass NTT: grid too large");
  const size_t smem = sizeof(uint32_t) * (size_t)BigPlan<BIG_R2>::PITCH * lanes;
  if (lanes == 32) {
    auto kern = ntt_big_strided_kernel<INV, BIG_R2, 32>;
    static bool attr = false;
    if (!attr) {
      B2Z_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr = true;
    }
    kern<<<(unsigned)grid, (32 << BIG_R2) / 32, smem, ctx->stream>>>(a);
  } else {
    auto kern = ntt_big_strided_kernel<INV, BIG_R2, 16>;
    static bool attr = false;
    if (!attr) {
      B2Z_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      attr = true;
    }
    kern<<<(unsigned)grid, (16 << BIG_R2) / 32, smem, ctx->stream>>>(a);
  }
  ctx->launches++;
  B2Z_CUDA(cudaGetLastError());
  return nullptr;
}

template <int R, int FIRST, bool EXPAND>
static b2z_err launch_big_rows_fwd_inst(b2z_ctx* ctx, BigArgs a) {
  using PL = BigPlan<R>;
  B2Z_TRY(big_tables<R>(ctx, false, a.gt));
  auto kern = ntt_big_rows_fwd_kernel<R, FIRST, EXPAND>;
  static bool attr = false;
  if (!attr) {
    B2Z_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PL::SMEM));
    attr = true;
  }
  const size_t rows = (size_t)a.cols << (a.logn - R);
  kern<<<(unsigned)(rows / PL::ROWS), PL::THREADS, PL::SMEM, ctx->stream>>>(a);
  ctx->launches++;
  B2Z_CUDA(cudaGetLastError());
  return nullptr;
}
template <int R>
static b2z_err launch_big_rows_inv_inst(b2z_ctx* ctx, BigArgs a) {
  using PL = BigPlan<R>;
  B2Z_TRY(big_tables<R>(ctx, true, a.gt));
  auto kern = ntt_big_rows_inv_kernel<R>;
  static bool attr = false;
  if (!attr) {
    B2Z_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PL::SMEM));
    attr = true;
  }
  const size_t rows = (size_t)a.cols << (a.logn - R);
  kern<<<(unsigned)(rows / PL::ROWS), PL::THREADS, PL::SMEM, ctx->stream>>>(a);
  ctx->launches++;
  B2Z_CUDA(cudaGetLastError());
  return nullptr;
}

// forward: out (2^logn per column) from in (2^(logn-2) per column when expanding; in == out otherwise)
static b2z_err launch_big_evaluate(b2z_ctx* ctx, uint32_t* out, const uint32_t* in, size_t count, unsigned logn, bool expand) {
  const unsigned r1 = logn - BIG_R2;
  // column chunks keep the grid of the strided pass below 2^31 CTAs and `cols` inside 32 bits
  const size_t max_cols = 4096;
  for (size_t c0 = 0; c0 < count; c0 += max_cols) {
    const size_t cols = count - c0 < max_cols ? count - c0 : max_cols;
    BigArgs a{};
    a.out = out + (c0 << logn);
    a.in = expand ? in + (c0 << (logn - 2)) : out + (c0 << logn);
    a.logn = logn;
    a.s_lo = 0;
    a.cols = (unsigned)cols;
#define B2Z_BIG_FWD(RR)                                                                    \
  case RR:                                                                                 \
    if (expand) B2Z_TRY((launch_big_rows_fwd_inst<RR, 3, true>(ctx, a)));                  \
    else B2Z_TRY((launch_big_rows_fwd_inst<RR, 1, false>(ctx, a)));                        \
    break;
    switch (r1) {
      B2Z_BIG_FWD(10) B2Z_BIG_FWD(11) B2Z_BIG_FWD(12) B2Z_BIG_FWD(13) B2Z_BIG_FWD(14)
      default: return set_error("two-pass NTT: unsupported size");
    }
#undef B2Z_BIG_FWD
    BigArgs b{};
    b.out = a.out;
    b.in = a.out;
    b.logn = logn;
    b.s_lo = r1;
    b.cols = (unsigned)cols;
    B2Z_TRY(launch_big_strided<false>(ctx, b));
  }
  return nullptr;
}

// inverse: io (2^logn per column), optionally reading the natural-order evaluations from `src`
static b2z_err launch_big_interpolate(b2z_ctx* ctx, uint32_t* io, const uint32_t* src, size_t count, unsigned logn, uint32_t norm,
                                      const uint32_t* scale_tab) {
  const unsigned r1 = logn - BIG_R2;
  const size_t max_cols = 4096;
  for (size_t c0 = 0; c0 < count; c0 += max_cols) {
    const size_t cols = count - c0 < max_cols ? count - c0 : max_cols;
    BigArgs b{};
    b.out = io + (c0 << logn);
    b.in = src ? src + (c0 << logn) : b.out;
    b.logn = logn;
    b.s_lo = r1;
    b.cols = (unsigned)cols;
    B2Z_TRY(launch_big_strided<true>(ctx, b));
    BigArgs a{};
    a.out = b.out;
    a.in = b.out;
    a.logn = logn;
    a.s_lo = 0;
    a.cols = (unsigned)cols;
    a.scale = norm;
    a.scale_tab = scale_tab;
    switch (r1) {
      case 10: B2Z_TRY(launch_big_rows_inv_inst<10>(ctx, a)); break;
      case 11: B2Z_TRY(launch_big_rows_inv_inst<11>(ctx, a)); break;
      case 12: B2Z_TRY(launch_big_rows_inv_inst<12>(ctx, a)); break;
      case 13: B2Z_TRY(launch_big_rows_inv_inst<13>(ctx, a)); break;
      case 14: B2Z_TRY(launch_big_rows_inv_inst<14>(ctx, a)); break;
      default: return set_error("two-pass NTT: unsupported size");
    }
  }
  return nullptr;
}

}  // namespace b2z
