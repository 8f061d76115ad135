// Polynomial / element-wise kernels of the segment prover: fri_fold, batch_evaluate_any, mix_poly_coeffs,
// eltwise_*, gather_sample, DEEP synthetic division, accumulate, and the batched query gather.
//
// Semantics: risc0-zkp 3.0.2 hal/cpu.rs (fri_fold, batch_evaluate_any, mix_poly_coeffs, eltwise_*, gather_sample),
// core/poly.rs poly_divide and prove/merkle.rs MerkleTreeProver::prove (rows a13-a16).
// All of these are HBM-bandwidth bound: one coalesced read of the inputs, ext-field arithmetic in registers.
#include <algorithm>

#include "common.h"

namespace b2z {

// ---------------- fri_fold ----------------
// out[k][idx] = sum_{i<16} mix^i * in[k'][rev4(i)*count + idx]  on component-major [4][.] buffers
__global__ void fri_fold_kernel(uint32_t* __restrict__ out, const uint32_t* __restrict__ in, Fp4 mix, size_t count) {
  size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= count) return;
  size_t in_count = count * 16;
  Fp4 tot = fp4_zero(), cur = fp4_one();
#pragma unroll
  for (unsigned i = 0; i < 16; i++) {
    size_t ri = (size_t)bit_rev(i, 4) * count + idx;
    Fp4 f = fp4_make(__ldg(in + ri), __ldg(in + in_count + ri), __ldg(in + 2 * in_count + ri), __ldg(in + 3 * in_count + ri));
    tot = fp4_add(tot, fp4_mul(cur, f));
    cur = fp4_mul(cur, mix);
  }
#pragma unroll
  for (int k = 0; k < 4; k++) out[k * count + idx] = tot.e[k];
}
b2z_err launch_fri_fold(b2z_ctx* ctx, uint32_t* out, const uint32_t* in, const Fp4& mix, size_t count) {
  if (!count) return nullptr;
  fri_fold_kernel<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(out, in, mix, count);
  ctx->launches++;
  B2Z_CUDA(cudaGetLastError());
  return nullptr;
}

// ---------------- batch_evaluate_any ----------------
// out[i] = sum_j coeffs[which[i]][j] * xs[i]^j (natural-order coefficients).
// Evaluations that share the same point x (all taps with the same `back`) share one table of powers x^j (ext, built
// once per distinct x, L2 resident); each evaluation is then a base-field x ext dot product: one coalesced read of the
// column, 4 modular multiplies per coefficient.  A block handles EA_GROUP evaluations of the same x over one chunk of
// coefficients so that every loaded power is reused EA_GROUP times from registers.
constexpr unsigned EA_THREADS = 256;
constexpr unsigned EA_PER_THREAD = 32;
constexpr unsigned EA_GROUP = 8;
constexpr unsigned EA_POW_RUN = 64;
__global__ void ea_powers_kernel(uint32_t* __restrict__ pw, Fp4 x, size_t n) {
  size_t b = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * EA_POW_RUN;
  if (b >= n) return;
  Fp4 cur = fp4_pow(x, b);
  for (size_t j = b; j < b + EA_POW_RUN && j < n; j++) {
    *reinterpret_cast<uint4*>(pw + 4 * j) = make_uint4(cur.e[0], cur.e[1], cur.e[2], cur.e[3]);
    cur = fp4_mul(cur, x);
  }
}
// group g: evaluations evs[g*EA_GROUP .. +EA_GROUP) (column ids in cols[], -1 padded), all at the point of table pw_of[g].
// A thread owns quads of 4 consecutive coefficients: one 128-bit load per column and four 128-bit loads of powers per quad.
__global__ void __launch_bounds__(EA_THREADS) ea_dot_kernel(const uint32_t* __restrict__ coeffs, size_t n,
                                                            const int* __restrict__ cols, const uint32_t* const* __restrict__ pw_of,
                                                            uint32_t* __restrict__ partial, unsigned chunks) {
  __shared__ Fp4 red[EA_THREADS];
  const unsigned g = blockIdx.y, chunk = blockIdx.x, t = threadIdx.x;
  const uint32_t* pw = pw_of[g];
  int col[EA_GROUP];
#pragma unroll
  for (unsigned e = 0; e < EA_GROUP; e++) col[e] = cols[g * EA_GROUP + e];
  Fp4 acc[EA_GROUP];
#pragma unroll
  for (unsigned e = 0; e < EA_GROUP; e++) acc[e] = fp4_zero();
  const size_t base = (size_t)chunk * EA_THREADS * EA_PER_THREAD;
  if (n >= 4) {
#pragma unroll 1
    for (unsigned k = 0; k < EA_PER_THREAD / 4; k++) {
      size_t j = base + 4 * ((size_t)t + (size_t)k * EA_THREADS);
      if (j >= n) break;
      uint4 p0 = *reinterpret_cast<const uint4*>(pw + 4 * j);
      uint4 p1 = *reinterpret_cast<const uint4*>(pw + 4 * j + 4);
      uint4 p2 = *reinterpret_cast<const uint4*>(pw + 4 * j + 8);
      uint4 p3 = *reinterpret_cast<const uint4*>(pw + 4 * j + 12);
      uint4 cv[EA_GROUP];
#pragma unroll
      for (unsigned e = 0; e < EA_GROUP; e++)
        cv[e] = col[e] >= 0 ? __ldg(reinterpret_cast<const uint4*>(coeffs + (size_t)col[e] * n + j)) : make_uint4(0, 0, 0, 0);
#pragma unroll
      for (unsigned e = 0; e < EA_GROUP; e++) {
        Fp4 a = acc[e];
        a.e[0] = fp_add(a.e[0], fp_dot4(p0.x, cv[e].x, p1.x, cv[e].y, p2.x, cv[e].z, p3.x, cv[e].w));
        a.e[1] = fp_add(a.e[1], fp_dot4(p0.y, cv[e].x, p1.y, cv[e].y, p2.y, cv[e].z, p3.y, cv[e].w));
        a.e[2] = fp_add(a.e[2], fp_dot4(p0.z, cv[e].x, p1.z, cv[e].y, p2.z, cv[e].z, p3.z, cv[e].w));
        a.e[3] = fp_add(a.e[3], fp_dot4(p0.w, cv[e].x, p1.w, cv[e].y, p2.w, cv[e].z, p3.w, cv[e].w));
        acc[e] = a;
      }
    }
  } else {  // tiny polynomials (n < 4): scalar path
    for (size_t j = base + t; j < n && j < base + (size_t)EA_THREADS * EA_PER_THREAD; j += EA_THREADS) {
      uint4 pv = *reinterpret_cast<const uint4*>(pw + 4 * j);
      for (unsigned e = 0; e < EA_GROUP; e++) {
        if (col[e] < 0) continue;
        uint32_t c = __ldg(coeffs + (size_t)col[e] * n + j);
        acc[e].e[0] = fp_add(acc[e].e[0], fp_mul(pv.x, c));
        acc[e].e[1] = fp_add(acc[e].e[1], fp_mul(pv.y, c));
        acc[e].e[2] = fp_add(acc[e].e[2], fp_mul(pv.z, c));
        acc[e].e[3] = fp_add(acc[e].e[3], fp_mul(pv.w, c));
      }
    }
  }
  // block reduction: warp shuffles first, then one shared-memory step per evaluation
  for (unsigned e = 0; e < EA_GROUP; e++) {
    Fp4 v = acc[e];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      Fp4 o = fp4_make(__shfl_down_sync(0xffffffffu, v.e[0], off), __shfl_down_sync(0xffffffffu, v.e[1], off),
                       __shfl_down_sync(0xffffffffu, v.e[2], off), __shfl_down_sync(0xffffffffu, v.e[3], off));
      v = fp4_add(v, o);
    }
    if ((t & 31) == 0) red[e * (EA_THREADS / 32) + (t >> 5)] = v;
  }
  __syncthreads();
  if (t < EA_GROUP) {
    Fp4 tot = fp4_zero();
    for (unsigned w = 0; w < EA_THREADS / 32; w++) tot = fp4_add(tot, red[t * (EA_THREADS / 32) + w]);
    uint32_t* o = partial + 4 * (((size_t)g * EA_GROUP + t) * chunks + chunk);
    o[0] = tot.e[0]; o[1] = tot.e[1]; o[2] = tot.e[2]; o[3] = tot.e[3];
  }
}
// out[dst[s]] = sum over chunks of partial[s]
__global__ void ea_reduce_kernel(const uint32_t* __restrict__ partial, const int* __restrict__ dst, uint32_t* __restrict__ out,
                                 unsigned chunks, size_t slots) {
  size_t s = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (s >= slots || dst[s] < 0) return;
  Fp4 tot = fp4_zero();
  for (unsigned c = 0; c < chunks; c++) {
    const uint32_t* q = partial + 4 * (s * chunks + c);
    tot = fp4_add(tot, fp4_make(q[0], q[1], q[2], q[3]));
  }
  uint32_t* o = out + 4 * (size_t)dst[s];
  o[0] = tot.e[0]; o[1] = tot.e[1]; o[2] = tot.e[2]; o[3] = tot.e[3];
}
b2z_err launch_evaluate_any_host(b2z_ctx* ctx, const uint32_t* coeffs, unsigned po2, const uint32_t* which_host, const Fp4* xs_host,
                                 size_t eval_count, uint32_t* out) {
  if (!eval_count) return nullptr;
  const size_t n = (size_t)1 << po2;
  // distinct points
  std::vector<Fp4> pts;
  std::vector<std::vector<size_t>> members;
  for (size_t i = 0; i < eval_count; i++) {
    size_t k = 0;
    for (; k < pts.size(); k++)
      if (fp4_eq(pts[k], xs_host[i])) break;
    if (k == pts.size()) {
      pts.push_back(xs_host[i]);
      members.emplace_back();
    }
    members[k].push_back(i);
  }
  // groups of EA_GROUP evaluations sharing a point
  std::vector<int> cols, dst;
  std::vector<size_t> group_pt;
  for (size_t k = 0; k < pts.size(); k++)
    for (size_t m = 0; m < members[k].size(); m += EA_GROUP) {
      group_pt.push_back(k);
      for (unsigned e = 0; e < EA_GROUP; e++) {
        bool ok = m + e < members[k].size();
        cols.push_back(ok ? (int)which_host[members[k][m + e]] : -1);
        dst.push_back(ok ? (int)members[k][m + e] : -1);
      }
    }
  const size_t groups = group_pt.size(), slots = groups * EA_GROUP;
  const size_t chunk_elems = (size_t)EA_THREADS * EA_PER_THREAD;
  const unsigned chunks = (unsigned)((n + chunk_elems - 1) / chunk_elems);
  B2Z_REQUIRE(groups <= 65535, "too many evaluation groups for one launch");
  uint32_t *pw = nullptr, *partial = nullptr;
  int *cols_dev = nullptr, *dst_dev = nullptr;
  const uint32_t** pw_of_dev = nullptr;
  B2Z_CUDA(cudaMallocAsync(&pw, 16 * n * pts.size(), ctx->stream));
  B2Z_CUDA(cudaMallocAsync(&partial, 16 * slots * chunks, ctx->stream));
  B2Z_CUDA(cudaMallocAsync(&cols_dev, 4 * slots, ctx->stream));
  B2Z_CUDA(cudaMallocAsync(&dst_dev, 4 * slots, ctx->stream));
  B2Z_CUDA(cudaMallocAsync(&pw_of_dev, 8 * groups, ctx->stream));
  std::vector<const uint32_t*> pw_of(groups);
  for (size_t g = 0; g < groups; g++) pw_of[g] = pw + 4 * n * group_pt[g];
  B2Z_CUDA(cudaMemcpyAsync(cols_dev, cols.data(), 4 * slots, cudaMemcpyHostToDevice, ctx->stream));
  B2Z_CUDA(cudaMemcpyAsync(dst_dev, dst.data(), 4 * slots, cudaMemcpyHostToDevice, ctx->stream));
  B2Z_CUDA(cudaMemcpyAsync(pw_of_dev, pw_of.data(), 8 * groups, cudaMemcpyHostToDevice, ctx->stream));
  size_t pthreads = (n + EA_POW_RUN - 1) / EA_POW_RUN;
  for (size_t k = 0; k < pts.size(); k++) {
    ea_powers_kernel<<<(unsigned)((pthreads + 127) / 128), 128, 0, ctx->stream>>>(pw + 4 * n * k, pts[k], n);
    ctx->launches++;
  }
  ea_dot_kernel<<<dim3(chunks, (unsigned)groups), EA_THREADS, 0, ctx->stream>>>(coeffs, n, cols_dev, pw_of_dev, partial, chunks);
  ea_reduce_kernel<<<(unsigned)((slots + 127) / 128), 128, 0, ctx->stream>>>(partial, dst_dev, out, chunks, slots);
  ctx->launches += 2;
  B2Z_CUDA(cudaGetLastError());
  B2Z_CUDA(cudaFreeAsync(pw, ctx->stream));
  B2Z_CUDA(cudaFreeAsync(partial, ctx->stream));
  B2Z_CUDA(cudaFreeAsync(cols_dev, ctx->stream));
  B2Z_CUDA(cudaFreeAsync(dst_dev, ctx->stream));
  B2Z_CUDA(cudaFreeAsync(pw_of_dev, ctx->stream));
  return nullptr;
}
// device-array form of the C ABI: fetch the (small) which / xs arrays and use the grouped implementation
b2z_err launch_evaluate_any(b2z_ctx* ctx, const uint32_t* coeffs, unsigned po2, const uint32_t* which, const uint32_t* xs,
                            size_t eval_count, uint32_t* out) {
  if (!eval_count) return nullptr;
  std::vector<uint32_t> w(eval_count);
  std::vector<Fp4> x(eval_count);
  B2Z_CUDA(cudaMemcpyAsync(w.data(), which, 4 * eval_count, cudaMemcpyDeviceToHost, ctx->stream));
  B2Z_CUDA(cudaMemcpyAsync(x.data(), xs, 16 * eval_count, cudaMemcpyDeviceToHost, ctx->stream));
  B2Z_CUDA(cudaStreamSynchronize(ctx->stream));
  return launch_evaluate_any_host(ctx, coeffs, po2, w.data(), x.data(), eval_count, out);
}

// ---------------- mix_poly_coeffs ----------------
// out[combos[i]][idx] += mix_start * mix^i * in[i][idx]   (out: AoS ext [n_combos+1][count])
// Up to MPC_MAX_COMBOS accumulators stay in registers; powers of mix come from a device table.
constexpr int MPC_MAX_COMBOS = 8;
__global__ void mix_pow_table_kernel(uint32_t* tab, Fp4 start, Fp4 mix, size_t n) {
  // tab[i] = start * mix^i ; one thread per 64 entries (pow for the first, then running product)
  size_t b = (blockIdx.x * (size_t)blockDim.x + threadIdx.x) * 64;
  if (b >= n) return;
  Fp4 cur = fp4_mul(start, fp4_pow(mix, b));
  for (size_t i = b; i < b + 64 && i < n; i++) {
    tab[4 * i] = cur.e[0]; tab[4 * i + 1] = cur.e[1]; tab[4 * i + 2] = cur.e[2]; tab[4 * i + 3] = cur.e[3];
    cur = fp4_mul(cur, mix);
  }
}
// Columns are visited in combo order (order[] = column indices sorted by combo id, seg[k]..seg[k+1] = the columns of the
// k-th combo of this launch), so the inner loop is branch-free and each accumulator is a plain register.
struct MpcSegs {
  uint32_t combo[MPC_MAX_COMBOS];    // output combo id of slot k
  uint32_t seg[MPC_MAX_COMBOS + 1];  // range of order[] belonging to slot k
  uint32_t n;                        // slots used
};
__global__ void __launch_bounds__(256) mix_poly_coeffs_kernel(uint32_t* __restrict__ out, const uint32_t* __restrict__ pows,
                                                              const uint32_t* __restrict__ in, const uint32_t* __restrict__ order,
                                                              size_t count, MpcSegs segs) {
  size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= count) return;
#pragma unroll 1
  for (unsigned k = 0; k < segs.n; k++) {
    Fp4 acc = fp4_zero();
    for (unsigned j = segs.seg[k]; j < segs.seg[k + 1]; j++) {
      unsigned i = order[j];
      uint32_t v = __ldg(in + (size_t)i * count + idx);
      const uint4 pw = *reinterpret_cast<const uint4*>(pows + 4 * (size_t)i);
      acc.e[0] = fp_add(acc.e[0], fp_mul(pw.x, v));
      acc.e[1] = fp_add(acc.e[1], fp_mul(pw.y, v));
      acc.e[2] = fp_add(acc.e[2], fp_mul(pw.z, v));
      acc.e[3] = fp_add(acc.e[3], fp_mul(pw.w, v));
    }
    uint4* o = reinterpret_cast<uint4*>(out + 4 * ((size_t)segs.combo[k] * count + idx));
    uint4 cur = *o;
    Fp4 r = fp4_add(fp4_make(cur.x, cur.y, cur.z, cur.w), acc);
    *o = make_uint4(r.e[0], r.e[1], r.e[2], r.e[3]);
  }
}
b2z_err launch_mix_poly_coeffs_host(b2z_ctx* ctx, uint32_t* out, const Fp4& mix_start, const Fp4& mix, const uint32_t* in,
                                    const uint32_t* combos_host, size_t input_size, size_t count) {
  if (!input_size || !count) return nullptr;
  // column order sorted by combo id (stable), split into launches of at most MPC_MAX_COMBOS distinct combos
  std::vector<uint32_t> order(input_size);
  for (size_t i = 0; i < input_size; i++) order[i] = (uint32_t)i;
  std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) { return combos_host[a] < combos_host[b]; });
  uint32_t *pows = nullptr, *order_dev = nullptr;
  B2Z_CUDA(cudaMallocAsync(&pows, 16 * input_size, ctx->stream));
  B2Z_CUDA(cudaMallocAsync(&order_dev, 4 * input_size, ctx->stream));
  B2Z_CUDA(cudaMemcpyAsync(order_dev, order.data(), 4 * input_size, cudaMemcpyHostToDevice, ctx->stream));
  size_t tthreads = (input_size + 63) / 64;
  mix_pow_table_kernel<<<(unsigned)((tthreads + 63) / 64), 64, 0, ctx->stream>>>(pows, mix_start, mix, input_size);
  ctx->launches++;
  size_t j = 0;
  while (j < input_size) {
    MpcSegs segs{};
    while (j < input_size && segs.n < MPC_MAX_COMBOS) {
      uint32_t cid = combos_host[order[j]];
      segs.combo[segs.n] = cid;
      segs.seg[segs.n] = (uint32_t)j;
      while (j < input_size && combos_host[order[j]] == cid) j++;
      segs.n++;
      segs.seg[segs.n] = (uint32_t)j;
    }
    mix_poly_coeffs_kernel<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(out, pows, in, order_dev, count, segs);
    ctx->launches++;
  }
  B2Z_CUDA(cudaGetLastError());
  B2Z_CUDA(cudaFreeAsync(pows, ctx->stream));
  B2Z_CUDA(cudaFreeAsync(order_dev, ctx->stream));
  return nullptr;
}
// device-array form of the C ABI: fetch the (small) combo-id array
b2z_err launch_mix_poly_coeffs(b2z_ctx* ctx, uint32_t* out, const Fp4& mix_start, const Fp4& mix, const uint32_t* in,
                               const uint32_t* combos_dev, size_t input_size, size_t count) {
  if (!input_size || !count) return nullptr;
  std::vector<uint32_t> ids(input_size);
  B2Z_CUDA(cudaMemcpyAsync(ids.data(), combos_dev, 4 * input_size, cudaMemcpyDeviceToHost, ctx->stream));
  B2Z_CUDA(cudaStreamSynchronize(ctx->stream));
  return launch_mix_poly_coeffs_host(ctx, out, mix_start, mix, in, ids.data(), input_size, count);
}

// ---------------- element-wise ----------------
__global__ void add_elem_kernel(uint32_t* out, const uint32_t* a, const uint32_t* b, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n) out[i] = fp_add(a[i], b[i]);
}
b2z_err launch_add_elem(b2z_ctx* ctx, uint32_t* out, const uint32_t* a, const uint32_t* b, size_t n) {
  if (!n) return nullptr;
  add_elem_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(out, a, b, n);
  ctx->launches++;
  B2Z_CUDA(cudaGetLastError());
  return nullptr;
}
// AoS ext input [to_add][count] -> component-major [4][count]
__global__ void sum_extelem_kernel(uint32_t* __restrict__ out, const uint32_t* __restrict__ in, size_t count, size_t to_add) {
  size_t idx = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (idx >= count) return;
  Fp4 tot = fp4_zero();
  for (size_t i = 0; i < to_add; i++) {
    uint4 v = *reinterpret_cast<const uint4*>(in + 4 * (i * count + idx));
    tot = fp4_add(tot, fp4_make(v.x, v.y, v.z, v.w));
  }
#pragma unroll
  for (int k = 0; k < 4; k++) out[k * count + idx] = tot.e[k];
}
b2z_err launch_sum_extelem(b2z_ctx* ctx, uint32_t* out, const uint32_t* in, size_t count, size_t to_add) {
  if (!count) return nullptr;
  sum_extelem_kernel<<<(unsigned)((count + 255) / 256), 256, 0, ctx->stream>>>(out, in, count, to_add);
  ctx->launches++;
  B2Z_CUDA(cudaGetLastError());
  return nullptr;
}
__global__ void zeroize_kernel(uint32_t* io, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n && io[i] == INVALID) io[i] = 0;
}
b2z_err launch_zeroize(b2z_ctx* ctx, uint32_t* io, size_t n) {
  if (!n) return nullptr;
  zeroize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(io, n);
  ctx->launches++;
  B2Z_CUDA(cudaGetLastError());
  return nullptr;
}
__global__ void gather_sample_kernel(uint32_t* dst, const uint32_t* src, size_t idx, size_t size, size_t stride) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < size) dst[i] = src[idx + i * stride];
}
b2z_err launch_gather_sample(b2z_ctx* ctx, uint32_t* dst, const uint32_t* src, size_t idx, size_t size, size_t stride) {
  if (!size) return nullptr;
  gather_sample_kernel<<<(unsigned)((size + 255) / 256), 256, 0, ctx->stream>>>(dst, src, idx, size, stride);
  ctx->launches++;
  B2Z_CUDA(cudaGetLastError());
  return nullptr;
}

// ---------------- DEEP synthetic division ----------------
// In-place division of the ext polynomial p (count AoS coefficients, natural order) by (X - z):
//   q[i] = sum_{j>i} p[j] z^(j-i-1)   (core/poly.rs poly_divide), remainder = sum_j p[j] z^j.
// Parallel form: segment s of L coefficients computes S_s = sum_{j in seg} p[j] z^(j - start_s) (Horner); the
// carry into segment s is C_s = sum_{s'>s} S_s' z^(L (s'-s-1)), a suffix scan with multiplier z^L done in two
// levels (within a block through shared memory, across blocks by a single-block kernel).
constexpr unsigned PD_L = 8;         // coefficients per thread
constexpr unsigned PD_THREADS = 256; // segments per block
__device__ __forceinline__ Fp4 ld_ext(const uint32_t* p) {
  uint4 v = *reinterpret_cast<const uint4*>(p);
  return fp4_make(v.x, v.y, v.z, v.w);
}
__device__ __forceinline__ void st_ext(uint32_t* p, const Fp4& v) { *reinterpret_cast<uint4*>(p) = make_uint4(v.e[0], v.e[1], v.e[2], v.e[3]); }

// suffix scan inside a block: on return sh[t] = sum_{t' >= t} S_t' zL^(t'-t)
__device__ __forceinline__ void block_suffix_scan(Fp4* sh, Fp4 zL, unsigned t, unsigned nthreads) {
  Fp4 mult = zL;
  for (unsigned d = 1; d < nthreads; d <<= 1) {
    Fp4 add = fp4_zero();
    if (t + d < nthreads) add = fp4_mul(sh[t + d], mult);
    __syncthreads();
    sh[t] = fp4_add(sh[t], add);
    __syncthreads();
    mult = fp4_mul(mult, mult);
  }
}
__global__ void __launch_bounds__(PD_THREADS) poly_divide_local_kernel(const uint32_t* __restrict__ p, size_t count, Fp4 z,
                                                                       uint32_t* __restrict__ block_sums) {
  __shared__ Fp4 sh[PD_THREADS];
  unsigned t = threadIdx.x;
  size_t start = ((size_t)blockIdx.x * PD_THREADS + t) * PD_L;
  Fp4 acc = fp4_zero();
#pragma unroll
  for (int k = PD_L - 1; k >= 0; k--) {
    size_t j = start + k;
    Fp4 c = j < count ? ld_ext(p + 4 * j) : fp4_zero();
    acc = fp4_add(fp4_mul(acc, z), c);
  }
  sh[t] = acc;
  __syncthreads();
  block_suffix_scan(sh, fp4_pow(z, PD_L), t, PD_THREADS);
  if (t == 0) st_ext(block_sums + 4 * (size_t)blockIdx.x, sh[0]);
}
// single block: carry[b] = sum_{b' > b} blocksum[b'] * zB^(b'-b-1), zB = z^(L*THREADS); remainder = carry[-1].
// Thread t owns the contiguous run of `per` block sums [t*per, (t+1)*per): local Horner, block-wide suffix scan with
// multiplier zB^per, then a second walk over the run to emit the per-block carries.
__global__ void __launch_bounds__(PD_THREADS) poly_divide_carry_kernel(const uint32_t* __restrict__ block_sums, uint32_t* __restrict__ carries,
                                                                       size_t nblocks, Fp4 zB, uint32_t* __restrict__ remainder) {
  __shared__ Fp4 sh[PD_THREADS];
  const unsigned t = threadIdx.x;
  const size_t per = (nblocks + PD_THREADS - 1) / PD_THREADS;
  const size_t b0 = (size_t)t * per;
  Fp4 acc = fp4_zero();  // sum_{b in run} blocksum[b] * zB^(b - b0)
  for (size_t k = per; k-- > 0;) {
    size_t b = b0 + k;
    Fp4 v = b < nblocks ? ld_ext(block_sums + 4 * b) : fp4_zero();
    acc = fp4_add(fp4_mul(acc, zB), v);
  }
  sh[t] = acc;
  __syncthreads();
  Fp4 zRun = fp4_pow(zB, per);
  block_suffix_scan(sh, zRun, t, PD_THREADS);
  // carry entering the top of this thread's run = suffix of the later runs
  Fp4 cur = (t + 1 < PD_THREADS) ? sh[t + 1] : fp4_zero();
  for (size_t k = per; k-- > 0;) {
    size_t b = b0 + k;
    if (b < nblocks) {
      st_ext(carries + 4 * b, cur);
      cur = fp4_add(fp4_mul(cur, zB), ld_ext(block_sums + 4 * b));
    } else {
      cur = fp4_mul(cur, zB);
    }
  }
  if (t == 0) st_ext(remainder, cur);
}
__global__ void __launch_bounds__(PD_THREADS) poly_divide_apply_kernel(uint32_t* __restrict__ p, size_t count, Fp4 z,
                                                                       const uint32_t* __restrict__ carries) {
  __shared__ Fp4 sh[PD_THREADS + 1];
  unsigned t = threadIdx.x;
  size_t start = ((size_t)blockIdx.x * PD_THREADS + t) * PD_L;
  Fp4 c[PD_L];
  Fp4 acc = fp4_zero();
#pragma unroll
  for (int k = PD_L - 1; k >= 0; k--) {
    size_t j = start + k;
    c[k] = j < count ? ld_ext(p + 4 * j) : fp4_zero();
    acc = fp4_add(fp4_mul(acc, z), c[k]);
  }
  sh[t] = acc;
  if (t == 0) sh[PD_THREADS] = fp4_zero();
  __syncthreads();
  Fp4 zL = fp4_pow(z, PD_L);
  block_suffix_scan(sh, zL, t, PD_THREADS);
  // carry into this thread's segment = (suffix of later segments in block) + block carry * zL^(segments after t)
  Fp4 later = (t + 1 < PD_THREADS) ? sh[t + 1] : fp4_zero();
  Fp4 bc = ld_ext(carries + 4 * (size_t)blockIdx.x);
  Fp4 cur = fp4_add(later, fp4_mul(bc, fp4_pow(zL, PD_THREADS - 1 - t)));
  // walk the segment downwards: q[j] = cur ; cur = cur*z + p[j]
#pragma unroll
  for (int k = PD_L - 1; k >= 0; k--) {
    size_t j = start + k;
    if (j < count) st_ext(p + 4 * j, cur);
    cur = fp4_add(fp4_mul(cur, z), c[k]);
  }
}
b2z_err launch_poly_divide(b2z_ctx* ctx, uint32_t* ext_coeffs, size_t count, const Fp4& z, uint32_t* remainder_dev) {
  if (!count) return nullptr;
  size_t per_block = (size_t)PD_L * PD_THREADS;
  size_t nblocks = (count + per_block - 1) / per_block;
  uint32_t* tmp = nullptr;
  B2Z_CUDA(cudaMallocAsync(&tmp, 16 * 2 * nblocks, ctx->stream));
  uint32_t* sums = tmp;
  uint32_t* carries = tmp + 4 * nblocks;
  poly_divide_local_kernel<<<(unsigned)nblocks, PD_THREADS, 0, ctx->stream>>>(ext_coeffs, count, z, sums);
  poly_divide_carry_kernel<<<1, PD_THREADS, 0, ctx->stream>>>(sums, carries, nblocks, fp4_pow(z, per_block), remainder_dev);
  poly_divide_apply_kernel<<<(unsigned)nblocks, PD_THREADS, 0, ctx->stream>>>(ext_coeffs, count, z, carries);
  ctx->launches += 3;
  B2Z_CUDA(cudaGetLastError());
  B2Z_CUDA(cudaFreeAsync(tmp, ctx->stream));
  return nullptr;
}
// p[i] -= sub[i] for the first n_sub ext coefficients (the interpolated U polynomial of a combo)
__global__ void sub_lowcoeffs_kernel(uint32_t* p, const uint32_t* sub, size_t n_sub) {
  size_t i = threadIdx.x;
  if (i < n_sub) st_ext(p + 4 * i, fp4_sub(ld_ext(p + 4 * i), ld_ext(sub + 4 * i)));
}
b2z_err launch_sub_lowcoeffs(b2z_ctx* ctx, uint32_t* ext_coeffs, const uint32_t* sub_dev, size_t n_sub) {
  if (!n_sub) return nullptr;
  B2Z_REQUIRE(n_sub <= 32, "U polynomial too long");
  sub_lowcoeffs_kernel<<<1, 32, 0, ctx->stream>>>(ext_coeffs, sub_dev, n_sub);
  ctx->launches++;
  B2Z_CUDA(cudaGetLastError());
  return nullptr;
}

// ---------------- accumulate (synthetic circuits) ----------------
// accumulator k: ext running sum over the active rows of 1 / (M_k + data[col_k][r]); noise in the last zk_rows rows.
// Stage 1: per-row inverse terms (ext) written to the accum columns; stage 2: per-accumulator inclusive scan.
__global__ void accum_terms_kernel(uint32_t* __restrict__ accum, const uint32_t* __restrict__ data, const uint32_t* __restrict__ cols,
                                   const uint32_t* __restrict__ mix, const uint32_t* __restrict__ noise, unsigned po2, size_t zk_rows) {
  size_t n = (size_t)1 << po2, active = n - zk_rows;
  size_t r = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t k = blockIdx.y;
  if (r >= n) return;
  uint32_t* a = accum + 4 * k * n;
  if (r >= active) {
#pragma unroll
    for (int j = 0; j < 4; j++) a[j * n + r] = noise[(4 * k + j) * zk_rows + (r - active)];
    return;
  }
  Fp4 m = fp4_make(mix[4 * k], mix[4 * k + 1], mix[4 * k + 2], mix[4 * k + 3]);
  Fp4 t = fp4_inv(fp4_add_fp(m, __ldg(data + (size_t)cols[k] * n + r)));
#pragma unroll
  for (int j = 0; j < 4; j++) a[j * n + r] = t.e[j];
}
constexpr unsigned SCAN_THREADS = 256;
constexpr unsigned SCAN_PER_THREAD = 16;
// level 1: block-local inclusive scan of ext terms over [0, active); block totals out
__global__ void __launch_bounds__(SCAN_THREADS) accum_scan_local_kernel(uint32_t* __restrict__ accum, unsigned po2, size_t active,
                                                                        uint32_t* __restrict__ block_tot, size_t nblocks) {
  __shared__ Fp4 sh[SCAN_THREADS];
  size_t n = (size_t)1 << po2;
  size_t k = blockIdx.y;
  uint32_t* a = accum + 4 * k * n;
  unsigned t = threadIdx.x;
  size_t start = ((size_t)blockIdx.x * SCAN_THREADS + t) * SCAN_PER_THREAD;
  Fp4 v[SCAN_PER_THREAD];
  Fp4 run = fp4_zero();
#pragma unroll
  for (unsigned i = 0; i < SCAN_PER_THREAD; i++) {
    size_t r = start + i;
    v[i] = r < active ? fp4_make(a[r], a[n + r], a[2 * n + r], a[3 * n + r]) : fp4_zero();
    run = fp4_add(run, v[i]);
    v[i] = run;
  }
  sh[t] = run;
  __syncthreads();
  for (unsigned d = 1; d < SCAN_THREADS; d <<= 1) {
    Fp4 add = fp4_zero();
    if (t >= d) add = sh[t - d];
    __syncthreads();
    sh[t] = fp4_add(sh[t], add);
    __syncthreads();
  }
  Fp4 prefix = t ? sh[t - 1] : fp4_zero();
#pragma unroll
  for (unsigned i = 0; i < SCAN_PER_THREAD; i++) {
    size_t r = start + i;
    if (r < active) {
      Fp4 o = fp4_add(prefix, v[i]);
      a[r] = o.e[0]; a[n + r] = o.e[1]; a[2 * n + r] = o.e[2]; a[3 * n + r] = o.e[3];
    }
  }
  if (t == SCAN_THREADS - 1) st_ext(block_tot + 4 * (k * nblocks + blockIdx.x), sh[t]);
}
__global__ void accum_scan_carry_kernel(uint32_t* __restrict__ block_tot, size_t nblocks) {
  // one thread per accumulator: exclusive scan of block totals in place
  size_t k = blockIdx.x;
  if (threadIdx.x != 0) return;
  Fp4 run = fp4_zero();
  for (size_t b = 0; b < nblocks; b++) {
    Fp4 v = ld_ext(block_tot + 4 * (k * nblocks + b));
    st_ext(block_tot + 4 * (k * nblocks + b), run);
    run = fp4_add(run, v);
  }
}
__global__ void __launch_bounds__(SCAN_THREADS) accum_scan_apply_kernel(uint32_t* __restrict__ accum, unsigned po2, size_t active,
                                                                        const uint32_t* __restrict__ block_tot, size_t nblocks) {
  size_t n = (size_t)1 << po2;
  size_t k = blockIdx.y;
  uint32_t* a = accum + 4 * k * n;
  Fp4 carry = ld_ext(block_tot + 4 * (k * nblocks + blockIdx.x));
  size_t base = (size_t)blockIdx.x * SCAN_THREADS * SCAN_PER_THREAD;
  for (unsigned i = threadIdx.x; i < SCAN_THREADS * SCAN_PER_THREAD; i += SCAN_THREADS) {
    size_t r = base + i;
    if (r < active) {
#pragma unroll
      for (int j = 0; j < 4; j++) a[j * n + r] = fp_add(a[j * n + r], carry.e[j]);
    }
  }
}
b2z_err launch_accum(b2z_ctx* ctx, uint32_t* accum, const uint32_t* data, const uint32_t* accum_cols_dev, size_t n_accum,
                     const uint32_t* mix_dev, const uint32_t* noise_dev, unsigned po2, size_t zk_rows) {
  if (!n_accum) return nullptr;
  size_t n = (size_t)1 << po2, active = n - zk_rows;
  accum_terms_kernel<<<dim3((unsigned)((n + 127) / 128), (unsigned)n_accum), 128, 0, ctx->stream>>>(accum, data, accum_cols_dev, mix_dev,
                                                                                                    noise_dev, po2, zk_rows);
  ctx->launches++;
  size_t per_block = (size_t)SCAN_THREADS * SCAN_PER_THREAD;
  size_t nblocks = (active + per_block - 1) / per_block;
  if (nblocks == 0) return nullptr;
  uint32_t* tot = nullptr;
  B2Z_CUDA(cudaMallocAsync(&tot, 16 * n_accum * nblocks, ctx->stream));
  accum_scan_local_kernel<<<dim3((unsigned)nblocks, (unsigned)n_accum), SCAN_THREADS, 0, ctx->stream>>>(accum, po2, active, tot, nblocks);
  accum_scan_carry_kernel<<<(unsigned)n_accum, 32, 0, ctx->stream>>>(tot, nblocks);
  accum_scan_apply_kernel<<<dim3((unsigned)nblocks, (unsigned)n_accum), SCAN_THREADS, 0, ctx->stream>>>(accum, po2, active, tot, nblocks);
  ctx->launches += 3;
  B2Z_CUDA(cudaGetLastError());
  B2Z_CUDA(cudaFreeAsync(tot, ctx->stream));
  return nullptr;
}

// ---------------- batched query gather ----------------
// For query q and tree t: write row `idx` (cols words, gathered from the column-major matrix) followed by the sibling
// digests from the leaf layer up to (not including) the top layer -- the byte layout MerkleTreeProver::prove emits.
__global__ void gather_queries_kernel(uint32_t* __restrict__ seal, const GatherTree* __restrict__ trees, size_t n_trees,
                                      const uint32_t* __restrict__ positions, size_t record_words) {
  size_t q = blockIdx.x, t = blockIdx.y;
  GatherTree tr = trees[t];
  uint32_t idx = positions[q] & (tr.rows - 1);
  uint32_t* o = seal + q * record_words + tr.seal_off;
  for (uint32_t c = threadIdx.x; c < tr.cols; c += blockDim.x) o[c] = tr.matrix[(size_t)c * tr.rows + idx];
  o += tr.cols;
  // path: node indices idx+rows, then halved while >= 2*top_size
  uint32_t node = idx + tr.rows;
  uint32_t level = 0;
  while (node >= 2 * tr.top_size) {
    if (threadIdx.x < 8) o[level * 8 + threadIdx.x] = tr.nodes[(size_t)(node ^ 1) * 8 + threadIdx.x];
    node >>= 1;
    level++;
  }
}
b2z_err launch_gather_queries(b2z_ctx* ctx, uint32_t* seal_dev, const GatherTree* trees_dev, size_t n_trees,
                              const uint32_t* positions_dev, size_t n_queries, size_t record_words) {
  gather_queries_kernel<<<dim3((unsigned)n_queries, (unsigned)n_trees), 128, 0, ctx->stream>>>(seal_dev, trees_dev, n_trees, positions_dev,
                                                                                              record_words);
  ctx->launches++;
  B2Z_CUDA(cudaGetLastError());
  return nullptr;
}

// ---------------- scatter / prefix_products (remaining trait Hal ops) ----------------
// Hal::scatter(into, index, offsets, values) [RECALL lo-med]: CSR-style -- `index` holds count + 1 row starts; cycle c owns the
// entries idx in [index[c], index[c+1]) and performs into[offsets[idx]] = values[idx] in that order (a later entry of
// the same cycle overwrites an earlier one; upstream runs the cycles in parallel, so cross-cycle duplicates are a race
// there as well).  One thread per cycle keeps exactly that order.
__global__ void scatter_kernel(uint32_t* __restrict__ into, const uint32_t* __restrict__ index, const uint32_t* __restrict__ offsets,
                               const uint32_t* __restrict__ values, size_t count) {
  size_t cycle = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (cycle >= count) return;
  for (uint32_t idx = index[cycle], end = index[cycle + 1]; idx < end; idx++) into[offsets[idx]] = values[idx];
}
b2z_err launch_scatter(b2z_ctx* ctx, uint32_t* into, const uint32_t* index, const uint32_t* offsets, const uint32_t* values, size_t count) {
  if (!count) return nullptr;
  scatter_kernel<<<(unsigned)((count + 127) / 128), 128, 0, ctx->stream>>>(into, index, offsets, values, count);
  ctx->launches++;
  B2Z_CUDA(cudaGetLastError());
  return nullptr;
}

// Hal::prefix_products(io) [RECALL med]: in-place inclusive prefix product of an AoS ext array, io[i] = io[0] * ... * io[i].
// Three launches: per-block inclusive scan (4 elements per thread, Hillis-Steele over the thread products in shared
// memory), one CTA turning the block totals into exclusive carries, and the carry multiplication.
constexpr unsigned PFX_THREADS = 256, PFX_PER_THREAD = 4, PFX_BLOCK = PFX_THREADS * PFX_PER_THREAD;
__global__ void __launch_bounds__(PFX_THREADS) prefix_products_local_kernel(uint32_t* __restrict__ io, size_t n, uint32_t* __restrict__ block_tot) {
  __shared__ Fp4 sh[PFX_THREADS];
  const unsigned t = threadIdx.x;
  const size_t base = (size_t)blockIdx.x * PFX_BLOCK + (size_t)t * PFX_PER_THREAD;
  Fp4 v[PFX_PER_THREAD];
  Fp4 run = fp4_one();
#pragma unroll
  for (unsigned k = 0; k < PFX_PER_THREAD; k++) {
    v[k] = base + k < n ? ld_ext(io + 4 * (base + k)) : fp4_one();
    run = fp4_mul(run, v[k]);
    v[k] = run;  // inclusive within the thread
  }
  sh[t] = run;
  __syncthreads();
  for (unsigned d = 1; d < PFX_THREADS; d <<= 1) {  // inclusive scan of the thread products (left factor first: order kept)
    Fp4 left = t >= d ? sh[t - d] : fp4_one();
    __syncthreads();
    if (t >= d) sh[t] = fp4_mul(left, sh[t]);
    __syncthreads();
  }
  const Fp4 excl = t ? sh[t - 1] : fp4_one();
#pragma unroll
  for (unsigned k = 0; k < PFX_PER_THREAD; k++)
    if (base + k < n) st_ext(io + 4 * (base + k), fp4_mul(excl, v[k]));
  if (t == PFX_THREADS - 1) st_ext(block_tot + 4 * blockIdx.x, sh[t]);
}
// block totals -> exclusive carries (in place), one CTA, chunks of PFX_THREADS totals with a running carry
__global__ void __launch_bounds__(PFX_THREADS) prefix_products_carry_kernel(uint32_t* __restrict__ block_tot, size_t nblocks) {
  __shared__ Fp4 sh[PFX_THREADS];
  __shared__ Fp4 carry;
  const unsigned t = threadIdx.x;
  if (t == 0) carry = fp4_one();
  __syncthreads();
  for (size_t b0 = 0; b0 < nblocks; b0 += PFX_THREADS) {
    const Fp4 own = b0 + t < nblocks ? ld_ext(block_tot + 4 * (b0 + t)) : fp4_one();
    sh[t] = own;
    __syncthreads();
    for (unsigned d = 1; d < PFX_THREADS; d <<= 1) {
      Fp4 left = t >= d ? sh[t - d] : fp4_one();
      __syncthreads();
      if (t >= d) sh[t] = fp4_mul(left, sh[t]);
      __syncthreads();
    }
    const Fp4 c0 = carry;
    const Fp4 excl = fp4_mul(c0, t ? sh[t - 1] : fp4_one());
    if (b0 + t < nblocks) st_ext(block_tot + 4 * (b0 + t), excl);
    __syncthreads();
    if (t == PFX_THREADS - 1) carry = fp4_mul(c0, sh[t]);
    __syncthreads();
  }
}
__global__ void __launch_bounds__(PFX_THREADS) prefix_products_apply_kernel(uint32_t* __restrict__ io, size_t n, const uint32_t* __restrict__ carries) {
  if (blockIdx.x == 0) return;  // carry of the first block is one
  const Fp4 c = ld_ext(carries + 4 * blockIdx.x);
  const size_t base = (size_t)blockIdx.x * PFX_BLOCK + (size_t)threadIdx.x * PFX_PER_THREAD;
#pragma unroll
  for (unsigned k = 0; k < PFX_PER_THREAD; k++)
    if (base + k < n) st_ext(io + 4 * (base + k), fp4_mul(c, ld_ext(io + 4 * (base + k))));
}
b2z_err launch_prefix_products(b2z_ctx* ctx, uint32_t* io, size_t n) {
  if (n <= 1) return nullptr;
  const size_t nblocks = (n + PFX_BLOCK - 1) / PFX_BLOCK;
  uint32_t* tot = nullptr;
  B2Z_CUDA(cudaMallocAsync(&tot, 16 * nblocks, ctx->stream));
  prefix_products_local_kernel<<<(unsigned)nblocks, PFX_THREADS, 0, ctx->stream>>>(io, n, tot);
  ctx->launches++;
  if (nblocks > 1) {
    prefix_products_carry_kernel<<<1, PFX_THREADS, 0, ctx->stream>>>(tot, nblocks);
    prefix_products_apply_kernel<<<(unsigned)nblocks, PFX_THREADS, 0, ctx->stream>>>(io, n, tot);
    ctx->launches += 2;
  }
  B2Z_CUDA(cudaGetLastError());
  B2Z_CUDA(cudaFreeAsync(tot, ctx->stream));
  return nullptr;
}

}  // namespace b2z
