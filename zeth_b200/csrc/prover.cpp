// Segment prover driver: the host-side orchestration of one STARK segment proof on one B200.
//
// Call sequence = risc0-circuit-rv32im 4.0.2 SegmentProver::prove over risc0-zkp 3.0.2 prove/{prover,poly_group,
// merkle,fri,write_iop}.rs: commit code/data groups, draw the accum mix,
// accumulate, commit accum, eval_check, commit the check group, DEEP evaluations + U coefficients, combos, FRI,
// 50 queries.  B200-first differences that keep the seal bit-identical:
//   * the whole pipeline stays on the device; the host only runs the Poseidon2 transcript on 32-byte roots;
//   * zk_shift is fused into the last iNTT pass; DEEP synthetic division runs on the GPU as a scan;
//   * all 50 query openings (rows + Merkle paths of every tree) are produced by ONE gather kernel, because the
//     query positions depend only on the transcript state, not on the opened data.
#include <nvtx3/nvToolsExt.h>  // header-only; ranges show up in Nsight tools, no-ops otherwise

#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>

#include "circuit.h"
#include "common.h"

namespace b2z {

static const char* PHASE_NAMES[] = {"commit_code", "commit_data", "accumulate", "commit_accum", "eval_check", "commit_check",
                                    "deep_evaluate", "deep_combos", "fri_commit", "fri_queries", "total"};
constexpr size_t N_PHASES = sizeof(PHASE_NAMES) / sizeof(PHASE_NAMES[0]);
const char* phase_name(size_t i) { return i < N_PHASES ? PHASE_NAMES[i] : ""; }

struct PhaseTimer {
  b2z_ctx* ctx;
  explicit PhaseTimer(b2z_ctx* c) : ctx(c) {
    if (ctx->phase_events.size() < N_PHASES) {
      ctx->phase_events.resize(N_PHASES);
      for (auto& pr : ctx->phase_events) {
        cudaEventCreate(&pr.first);
        cudaEventCreate(&pr.second);
      }
    }
  }
  void begin(size_t i) {
    nvtxRangePushA(PHASE_NAMES[i]);
    cudaEventRecord(ctx->phase_events[i].first, ctx->stream);
  }
  void end(size_t i) {
    cudaEventRecord(ctx->phase_events[i].second, ctx->stream);
    nvtxRangePop();
  }
  void collect() {
    ctx->phase_ms.assign(N_PHASES, 0.0);
    for (size_t i = 0; i < N_PHASES; i++) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, ctx->phase_events[i].first, ctx->phase_events[i].second) == cudaSuccess) ctx->phase_ms[i] = ms;
    }
  }
};

struct DevMem {  // stream-ordered allocations released together
  b2z_ctx* ctx;
  std::vector<void*> ptrs;
  explicit DevMem(b2z_ctx* c) : ctx(c) {}
  ~DevMem() {
    for (void* p : ptrs) cudaFreeAsync(p, ctx->stream);
  }
  b2z_err alloc(uint32_t** out, size_t words) {
    void* p = nullptr;
    B2Z_CUDA(cudaMallocAsync(&p, sizeof(uint32_t) * (words ? words : 1), ctx->stream));
    ptrs.push_back(p);
    *out = (uint32_t*)p;
    return nullptr;
  }
  void release(void* p) {
    for (auto& q : ptrs)
      if (q == p) {
        cudaFreeAsync(p, ctx->stream);
        q = ptrs.back();
        ptrs.pop_back();
        return;
      }
  }
};

struct MerkleParams {
  size_t rows, cols, layers, top_size;
  MerkleParams(size_t r, size_t c) : rows(r), cols(c) {
    layers = ilog2(r);
    top_size = (size_t)1 << pp_merkle_top_layer(layers, B2Z_QUERIES);  // [RECALL med] rule lives in protocol_params.h
  }
  size_t path_words() const { return 8 * (layers - ilog2(top_size)); }
};

// interior layers over leaves already sitting in nodes[rows .. 2 rows)
static b2z_err merkle_fold_layers(b2z_ctx* ctx, uint32_t* nodes, size_t rows) {
  // wide layers: one launch each; the top of the tree (<= 512 parents per layer) in a single launch
  constexpr size_t TOP = 512;
  size_t i = ilog2(rows);
  while (i-- > 0) {
    size_t out = (size_t)1 << i;
    if (out <= TOP) {
      B2Z_TRY(launch_hash_fold_top(ctx, nodes, out));
      break;
    }
    B2Z_TRY(launch_hash_fold(ctx, nodes, 2 * out, out));
  }
  return nullptr;
}

b2z_err merkle_build(b2z_ctx* ctx, uint32_t* nodes, const uint32_t* matrix, size_t rows, size_t cols) {
  B2Z_TRY(launch_hash_rows(ctx, nodes + 8 * rows, matrix, rows, cols));
  return merkle_fold_layers(ctx, nodes, rows);
}

struct PolyGroupDev {
  uint32_t *coeffs = nullptr, *evaluated = nullptr, *nodes = nullptr;
  size_t count = 0, n = 0;
};

struct Seal {  // pinned host image of the seal, filled in write order
  uint32_t* w;
  size_t pos = 0, cap;
};

static b2z_err ensure_pinned(b2z_ctx* ctx, size_t words) {
  if (ctx->pinned_words >= words) return nullptr;
  if (ctx->pinned) cudaFreeHost(ctx->pinned);
  ctx->pinned = nullptr;
  ctx->pinned_words = 0;
  B2Z_CUDA(cudaMallocHost(&ctx->pinned, sizeof(uint32_t) * words));
  ctx->pinned_words = words;
  return nullptr;
}

struct PolyGroupDev;
struct Seal;
// PolyGroup::new + merkle.commit: coeffs (bit-reversed, shifted) -> LDE, natural-order coeffs, Merkle tree, top layer + root
static b2z_err poly_group_finish(b2z_ctx* ctx, DevMem& mem, PolyGroupDev& pg, unsigned po2, Seal& seal, HostRng& rng, bool leaves_done = false);
static b2z_err poly_group_commit(b2z_ctx* ctx, DevMem& mem, PolyGroupDev& pg, uint32_t* coeffs, size_t count, unsigned po2, Seal& seal,
                                 HostRng& rng) {
  size_t n = (size_t)1 << po2, domain = 4 * n;
  pg.coeffs = coeffs;
  pg.count = count;
  pg.n = n;
  B2Z_TRY(mem.alloc(&pg.evaluated, count * domain));
  B2Z_TRY(launch_evaluate(ctx, pg.evaluated, coeffs, count, po2 + 2, 2));
  return poly_group_finish(ctx, mem, pg, po2, seal, rng);
}
// natural-order coefficients, Merkle tree over the LDE, top layer -> seal, root -> transcript
// (`leaves_done`: the caller already hashed the leaves chunk by chunk into pg.nodes)
static b2z_err poly_group_finish(b2z_ctx* ctx, DevMem& mem, PolyGroupDev& pg, unsigned po2, Seal& seal, HostRng& rng, bool leaves_done) {
  size_t n = (size_t)1 << po2, domain = 4 * n;
  size_t count = pg.count;
  uint32_t* coeffs = pg.coeffs;
  B2Z_TRY(launch_bit_reverse(ctx, coeffs, count, po2));
  if (leaves_done) {
    B2Z_TRY(merkle_fold_layers(ctx, pg.nodes, domain));
  } else {
    B2Z_TRY(mem.alloc(&pg.nodes, 2 * domain * 8));
    B2Z_TRY(merkle_build(ctx, pg.nodes, pg.evaluated, domain, count));
  }
  MerkleParams mp(domain, count);
  B2Z_REQUIRE(seal.pos + 8 * mp.top_size <= seal.cap, "seal overflow");
  B2Z_CUDA(cudaMemcpyAsync(seal.w + seal.pos, pg.nodes + 8 * mp.top_size, 32 * mp.top_size, cudaMemcpyDeviceToHost, ctx->stream));
  uint32_t* root = seal.w + seal.cap;  // scratch after the seal image (ensure_pinned reserves 64 extra words)
  B2Z_CUDA(cudaMemcpyAsync(root, pg.nodes + 8, 32, cudaMemcpyDeviceToHost, ctx->stream));
  B2Z_CUDA(b2z::stream_wait(ctx, ctx->stream));
  seal.pos += 8 * mp.top_size;
  rng.mix(root);
  return nullptr;
}

// Optional readiness of a host-uploaded witness: the upload runs on ctx->copy_stream in column chunks, the commit
// consumes each chunk as soon as its event fires: iNTT, LDE and -- when the chunk is a whole number of sponge blocks --
// the leaf hashing of its columns (launch_hash_rows_chunk carries the sponge capacity in the leaf slots), so the
// upload of chunk k+1 hides behind all the work on chunk k and only the interior Merkle layers remain at the end.
struct UploadGate;  // host-side "this event has been recorded" latch (pageable sources are staged by helper threads)
struct WitnessReady {
  cudaEvent_t noise = nullptr;
  std::vector<cudaEvent_t> code_chunks;  // one event per `code_chunk_cols` code columns (small: the first compute starts early)
  std::vector<cudaEvent_t> data_chunks;  // one event per `chunk_cols` data columns
  size_t chunk_cols = 0, code_chunk_cols = 0;
  // cudaStreamWaitEvent on an event that has not been recorded yet is a no-op, so when the copies are issued by the staging
  // threads the compute thread first blocks (on the host) until the chunk's event exists.  gate == nullptr: all recorded.
  UploadGate* gate = nullptr;
  size_t code_gate0 = 0, data_gate0 = 0, noise_gate = 0;  // gate indices of code chunk 0, data chunk 0 and the noise block
};
b2z_err upload_gate_wait(UploadGate* g, size_t idx);

// prover.rs make_coeffs + commit_group (PolyGroup::new inlined so that the per-chunk work can start early)
static b2z_err commit_group(b2z_ctx* ctx, DevMem& mem, PolyGroupDev& pg, const uint32_t* witness, size_t count, unsigned po2, Seal& seal,
                            HostRng& rng, const std::vector<cudaEvent_t>* chunk_events = nullptr, size_t chunk_cols = 0,
                            cudaEvent_t whole_event = nullptr, UploadGate* gate = nullptr, size_t gate0 = 0) {
  size_t n = (size_t)1 << po2, domain = 4 * n;
  uint32_t* coeffs = nullptr;
  B2Z_TRY(mem.alloc(&coeffs, count * n));
  pg.coeffs = coeffs;
  pg.count = count;
  pg.n = n;
  B2Z_TRY(mem.alloc(&pg.evaluated, count * domain));
  if (whole_event) B2Z_CUDA(cudaStreamWaitEvent(ctx->stream, whole_event, 0));
  if (!chunk_events || chunk_cols == 0) chunk_cols = count;
  const bool hash_chunks = chunk_cols < count && chunk_cols % 16 == 0;
  if (hash_chunks) B2Z_TRY(mem.alloc(&pg.nodes, 2 * domain * 8));
  for (size_t c0 = 0, k = 0; c0 < count; c0 += chunk_cols, k++) {
    size_t cols = count - c0 < chunk_cols ? count - c0 : chunk_cols;
    if (chunk_events && k < chunk_events->size()) {
      if (gate) B2Z_TRY(upload_gate_wait(gate, gate0 + k));
      B2Z_CUDA(cudaStreamWaitEvent(ctx->stream, (*chunk_events)[k], 0));
    }
    // the first inverse pass reads the witness directly and writes the coefficient buffer (no separate copy)
    B2Z_TRY(launch_interpolate(ctx, coeffs + c0 * n, cols, po2, /*fuse_zk_shift=*/true, witness + c0 * n));
    B2Z_TRY(launch_evaluate(ctx, pg.evaluated + c0 * domain, coeffs + c0 * n, cols, po2 + 2, 2));
    if (hash_chunks)
      B2Z_TRY(launch_hash_rows_chunk(ctx, pg.nodes + 8 * domain, pg.evaluated + c0 * domain, domain, cols, /*carry_in=*/c0 != 0,
                                     /*last=*/c0 + cols == count));
  }
  return poly_group_finish(ctx, mem, pg, po2, seal, rng, hash_chunks);
}

// unique interpolant through (x_i, fx_i) -> monomial coefficients (core/poly.rs poly_interpolate)
static void poly_interpolate(Fp4* out, const Fp4* x, const Fp4* fx, size_t size) {
  Fp4 dd[16], basis[17];
  for (size_t i = 0; i < size; i++) dd[i] = fx[i];
  for (size_t j = 1; j < size; j++)
    for (size_t i = size - 1; i >= j; i--) dd[i] = fp4_mul(fp4_sub(dd[i], dd[i - 1]), fp4_inv(fp4_sub(x[i], x[i - j])));
  for (size_t i = 0; i < size; i++) out[i] = fp4_zero();
  basis[0] = fp4_one();
  size_t blen = 1;
  for (size_t j = 0; j < size; j++) {
    for (size_t i = 0; i < blen; i++) out[i] = fp4_add(out[i], fp4_mul(dd[j], basis[i]));
    basis[blen] = fp4_zero();
    for (size_t i = blen; i > 0; i--) basis[i] = fp4_sub(basis[i - 1], fp4_mul(basis[i], x[j]));
    basis[0] = fp4_neg(fp4_mul(basis[0], x[j]));
    blen++;
  }
}

static void hash_info(const uint8_t* info, uint32_t* dg) {
  uint32_t e[16];
  for (int i = 0; i < 16; i++) e[i] = fp_from_u32(info[i]);
  host_hash_elem_slice(e, 16, dg);
}

b2z_err eval_check_launch(b2z_ctx* ctx, const b2z_circuit* c, uint32_t* check, const uint32_t* const groups[3], const uint32_t* globals_host,
                          const uint32_t* mix_host, const Fp4& poly_mix, unsigned po2);
b2z_err mix_pow_table(b2z_ctx* ctx, uint32_t* tab, const Fp4& start, const Fp4& mix, size_t n);

b2z_err prove_segment_device(b2z_ctx* ctx, const b2z_circuit* c, unsigned po2, const uint32_t* code_dev, const uint32_t* data_dev,
                             const uint32_t* globals_host, const uint32_t* accum_noise_dev, size_t zk_rows, uint32_t* seal_host,
                             size_t* seal_words_out, const WitnessReady* ready) {
  B2Z_REQUIRE(po2 >= 1 && po2 <= 24, "po2 out of range");
  B2Z_REQUIRE(c->library != nullptr, "circuit kernels not loaded");
  const size_t n = (size_t)1 << po2, domain = 4 * n;
  B2Z_REQUIRE(zk_rows < n, "zk_rows must be smaller than the trace");
  const size_t wa = c->group_size[0], wc = c->group_size[1], wd = c->group_size[2];
  const size_t gsz = c->globals_size, msz = c->mix_size;
  const size_t seal_words = circuit_seal_words(c, po2);
  B2Z_TRY(ensure_pinned(ctx, seal_words + 64 + 4 * (c->n_taps + 64) + 4 * (c->tot_combo_backs + 8) + 16));
  Seal seal{ctx->pinned, 0, seal_words};
  uint32_t* scratch = ctx->pinned + seal_words + 8;  // [root(8)] [eval_u ...] [DEEP remainders]
  uint32_t* rem_host = ctx->pinned + seal_words + 64 + 4 * (c->n_taps + 64);
  size_t n_rem = 0;
  DevMem mem(ctx);
  PhaseTimer tm(ctx);
  HostRng rng;
  uint32_t dg[8];
  tm.begin(10);

  hash_info(reinterpret_cast<const uint8_t*>(PP_PROOF_SYSTEM_INFO), dg);  // [RECALL med] protocol_params.h
  rng.mix(dg);
  hash_info(c->info, dg);
  rng.mix(dg);
  // header = globals || po2
  memcpy(seal.w, globals_host, 4 * gsz);
  seal.w[gsz] = fp_from_u32(po2);
  host_hash_elem_slice(seal.w, gsz + 1, dg);
  rng.mix(dg);
  seal.pos = gsz + 1;
  uint32_t* globals_dev = nullptr;
  B2Z_TRY(mem.alloc(&globals_dev, gsz + 1));
  B2Z_CUDA(cudaMemcpyAsync(globals_dev, seal.w, 4 * gsz, cudaMemcpyHostToDevice, ctx->stream));

  PolyGroupDev pg[3], check_pg;
  tm.begin(0);
  B2Z_TRY(commit_group(ctx, mem, pg[B2Z_GROUP_CODE], code_dev, wc, po2, seal, rng, ready ? &ready->code_chunks : nullptr,
                       ready ? ready->code_chunk_cols : 0, nullptr, ready ? ready->gate : nullptr, ready ? ready->code_gate0 : 0));
  tm.end(0);
  tm.begin(1);
  B2Z_TRY(commit_group(ctx, mem, pg[B2Z_GROUP_DATA], data_dev, wd, po2, seal, rng, ready ? &ready->data_chunks : nullptr,
                       ready ? ready->chunk_cols : 0, nullptr, ready ? ready->gate : nullptr, ready ? ready->data_gate0 : 0));
  tm.end(1);

  std::vector<uint32_t> mix_host(msz + 1);
  for (size_t i = 0; i < msz; i++) mix_host[i] = rng.random_elem();
  uint32_t* mix_dev = nullptr;
  B2Z_TRY(mem.alloc(&mix_dev, msz + 1));
  B2Z_CUDA(cudaMemcpyAsync(mix_dev, mix_host.data(), 4 * msz, cudaMemcpyHostToDevice, ctx->stream));
  tm.begin(2);
  uint32_t* accum = nullptr;
  B2Z_TRY(mem.alloc(&accum, wa * n));
  if (ready && ready->noise) {
    if (ready->gate) B2Z_TRY(upload_gate_wait(ready->gate, ready->noise_gate));
    B2Z_CUDA(cudaStreamWaitEvent(ctx->stream, ready->noise, 0));
  }
  B2Z_TRY(launch_accum(ctx, accum, data_dev, c->accum_col_dev, c->accum_col.size(), mix_dev, accum_noise_dev, po2, zk_rows));
  tm.end(2);
  tm.begin(3);
  {
    // the accum witness buffer itself becomes the coefficient buffer (no extra copy)
    B2Z_TRY(launch_interpolate(ctx, accum, wa, po2, true));
    B2Z_TRY(poly_group_commit(ctx, mem, pg[B2Z_GROUP_ACCUM], accum, wa, po2, seal, rng));
  }
  tm.end(3);

  // ---- finalize ----
  Fp4 poly_mix = rng.random_ext();
  uint32_t* check = nullptr;
  B2Z_TRY(mem.alloc(&check, 4 * domain));
  tm.begin(4);
  {
    const uint32_t* groups[3] = {pg[0].evaluated, pg[1].evaluated, pg[2].evaluated};
    B2Z_TRY(eval_check_launch(ctx, c, check, groups, globals_host, mix_host.data(), poly_mix, po2));
  }
  tm.end(4);
  tm.begin(5);
  B2Z_TRY(launch_interpolate(ctx, check, 4, po2 + 2, false));
  // Satisfiability check (upstream only debug-asserts; an unsatisfied witness would otherwise yield a well-formed but
  // invalid seal): constraints have degree <= 5 in taps of degree < N, so C(x) / (x^N - 1) has degree <= 4N - 5 and the four
  // highest coefficients of every component of the check polynomial are zero.  If some constraint fails on the trace, C is
  // not divisible by x^N - 1 and the interpolant of its coset values has full degree (with overwhelming probability).
  // The 16 words ride along with the next transcript round trip.
  uint32_t* top_host = rem_host + 4 * (c->tot_combo_backs + 2);
  for (unsigned t = 0; t < 4; t++) {
    const size_t slot = bit_rev((uint32_t)(domain - 1 - t), po2 + 2);  // coefficients are in bit-reversed order
    B2Z_CUDA(cudaMemcpy2DAsync(top_host + 4 * t, 4, check + slot, 4 * domain, 4, 4, cudaMemcpyDeviceToHost, ctx->stream));
  }
  // bit-reversed 4N coefficients per component = 4 quarters of N: the 16 x N check group, no data movement
  B2Z_TRY(poly_group_commit(ctx, mem, check_pg, check, B2Z_CHECK_SIZE, po2, seal, rng));
  for (unsigned i = 0; i < 16; i++)
    if (top_host[i] != 0)
      return set_error("witness does not satisfy the circuit: the check polynomial has degree > 4N - 5 (some constraint fails on the trace)");
  tm.end(5);

  tm.begin(6);
  Fp4 z = rng.random_ext();
  uint32_t back_one = rou_rev(po2);
  const size_t n_taps = c->n_taps, n_regs = c->reg_group.size(), n_combos = c->n_combos;
  const size_t n_evals = n_taps + B2Z_CHECK_SIZE;
  std::vector<Fp4> all_xs(n_evals);
  std::vector<uint32_t> which(n_evals);
  Fp4 z_pow = fp4_pow(z, B2Z_EXT_SIZE);
  {
    std::vector<uint32_t> back_pows;
    for (size_t t = 0; t < n_taps; t++) {
      uint32_t b = c->tap_back[t];
      while (back_pows.size() <= b) back_pows.push_back(back_pows.empty() ? MONT_ONE : fp_mul(back_pows.back(), back_one));
      which[t] = c->tap_offset[t];
      all_xs[t] = fp4_mul_fp(z, back_pows[b]);
    }
    for (size_t i = 0; i < B2Z_CHECK_SIZE; i++) {
      which[n_taps + i] = (uint32_t)i;
      all_xs[n_taps + i] = z_pow;
    }
  }
  uint32_t* evals_dev = nullptr;
  B2Z_TRY(mem.alloc(&evals_dev, 4 * n_evals));
  {
    size_t pos = 0;
    for (int g = 0; g < 3; g++) {
      size_t start = pos;
      while (pos < n_taps && c->tap_group[pos] == (uint32_t)g) pos++;
      B2Z_TRY(launch_evaluate_any_host(ctx, pg[g].coeffs, po2, which.data() + start, all_xs.data() + start, pos - start, evals_dev + 4 * start));
    }
    B2Z_TRY(launch_evaluate_any_host(ctx, check_pg.coeffs, po2, which.data() + n_taps, all_xs.data() + n_taps, B2Z_CHECK_SIZE,
                                     evals_dev + 4 * n_taps));
  }
  Fp4* eval_u = reinterpret_cast<Fp4*>(scratch + 8);
  B2Z_CUDA(cudaMemcpyAsync(eval_u, evals_dev, 16 * n_evals, cudaMemcpyDeviceToHost, ctx->stream));
  B2Z_CUDA(b2z::stream_wait(ctx, ctx->stream));
  B2Z_REQUIRE(seal.pos + 4 * n_evals <= seal.cap, "seal overflow");
  Fp4* coeff_u = reinterpret_cast<Fp4*>(seal.w + seal.pos);
  for (size_t r = 0; r < n_regs; r++) {
    size_t p0 = c->reg_tap_pos[r];
    B2Z_REQUIRE(c->reg_size[r] <= 16, "register with more than 16 taps");
    poly_interpolate(coeff_u + p0, all_xs.data() + p0, eval_u + p0, c->reg_size[r]);
  }
  for (size_t i = 0; i < B2Z_CHECK_SIZE; i++) coeff_u[n_taps + i] = eval_u[n_taps + i];
  seal.pos += 4 * n_evals;
  host_hash_elem_slice(reinterpret_cast<uint32_t*>(coeff_u), 4 * n_evals, dg);
  rng.mix(dg);
  tm.end(6);

  tm.begin(7);
  Fp4 mixv = rng.random_ext();
  uint32_t* combos = nullptr;
  B2Z_TRY(mem.alloc(&combos, 4 * n * (n_combos + 1)));
  B2Z_CUDA(cudaMemsetAsync(combos, 0, 16 * n * (n_combos + 1), ctx->stream));
  {
    // combo id of every column (host side: the launcher sorts the columns by combo)
    std::vector<uint32_t> ids(n_regs + B2Z_CHECK_SIZE);
    for (size_t r = 0; r < n_regs; r++) ids[r] = c->reg_combo[r];
    for (size_t i = 0; i < B2Z_CHECK_SIZE; i++) ids[n_regs + i] = (uint32_t)n_combos;
    Fp4 cur_mix = fp4_one();
    size_t r = 0;
    for (int g = 0; g < 3; g++) {
      size_t gs = c->group_size[g];
      B2Z_TRY(launch_mix_poly_coeffs_host(ctx, combos, cur_mix, mixv, pg[g].coeffs, ids.data() + r, gs, n));
      cur_mix = fp4_mul(cur_mix, fp4_pow(mixv, gs));
      r += gs;
    }
    B2Z_TRY(launch_mix_poly_coeffs_host(ctx, combos, cur_mix, mixv, check_pg.coeffs, ids.data() + n_regs, B2Z_CHECK_SIZE, n));
  }
  {
    // U polynomials per combo (same sums the verifier forms as combo_u), subtracted from the low coefficients
    std::vector<Fp4> combo_u(c->tot_combo_backs + 1, fp4_zero());
    Fp4 cur = fp4_one();
    size_t cur_pos = 0;
    for (size_t r = 0; r < n_regs; r++) {
      for (size_t i = 0; i < c->reg_size[r]; i++) {
        Fp4& dst = combo_u[c->combo_begin[c->reg_combo[r]] + i];
        dst = fp4_add(dst, fp4_mul(cur, coeff_u[cur_pos + i]));
      }
      cur = fp4_mul(cur, mixv);
      cur_pos += c->reg_size[r];
    }
    for (size_t i = 0; i < B2Z_CHECK_SIZE; i++) {
      combo_u[c->tot_combo_backs] = fp4_add(combo_u[c->tot_combo_backs], fp4_mul(cur, coeff_u[cur_pos]));
      cur_pos++;
      cur = fp4_mul(cur, mixv);
    }
    uint32_t *cu_dev = nullptr, *rem_dev = nullptr;
    B2Z_TRY(mem.alloc(&cu_dev, 4 * combo_u.size()));
    B2Z_TRY(mem.alloc(&rem_dev, 4 * (c->tot_combo_backs + 2)));
    B2Z_CUDA(cudaMemcpyAsync(cu_dev, combo_u.data(), 16 * combo_u.size(), cudaMemcpyHostToDevice, ctx->stream));
    std::vector<uint32_t> back_pows;
    size_t nrem = 0;
    for (size_t k = 0; k < n_combos; k++) {
      size_t b0 = c->combo_begin[k], b1 = c->combo_begin[k + 1];
      B2Z_TRY(launch_sub_lowcoeffs(ctx, combos + 4 * k * n, cu_dev + 4 * b0, b1 - b0));
      for (size_t j = b0; j < b1; j++) {
        uint32_t b = c->combo_backs[j];
        while (back_pows.size() <= b) back_pows.push_back(back_pows.empty() ? MONT_ONE : fp_mul(back_pows.back(), back_one));
        B2Z_TRY(launch_poly_divide(ctx, combos + 4 * k * n, n, fp4_mul_fp(z, back_pows[b]), rem_dev + 4 * nrem++));
      }
    }
    B2Z_TRY(launch_sub_lowcoeffs(ctx, combos + 4 * n_combos * n, cu_dev + 4 * c->tot_combo_backs, 1));
    B2Z_TRY(launch_poly_divide(ctx, combos + 4 * n_combos * n, n, z_pow, rem_dev + 4 * nrem++));
    // every division must be exact (prover.rs asserts a zero remainder): the remainders come back with the next transcript
    // round trip and an unsatisfied witness fails the segment instead of yielding a garbage seal
    n_rem = nrem;
    B2Z_CUDA(cudaMemcpyAsync(rem_host, rem_dev, 16 * n_rem, cudaMemcpyDeviceToHost, ctx->stream));
  }
  auto check_remainders = [&]() -> b2z_err {
    for (size_t i = 0; i < 4 * n_rem; i++)
      if (rem_host[i] != 0)
        return set_error("witness does not satisfy the circuit: non-zero remainder in DEEP division " + std::to_string(i / 4) +
                         " (constraint violated, or code/data/accum columns inconsistent with the transcript)");
    n_rem = 0;
    return nullptr;
  };
  uint32_t* final_coeffs = nullptr;
  B2Z_TRY(mem.alloc(&final_coeffs, 4 * n));
  B2Z_TRY(launch_sum_extelem(ctx, final_coeffs, combos, n, n_combos + 1));
  B2Z_TRY(launch_bit_reverse(ctx, final_coeffs, 4, po2));
  tm.end(7);

  // ---- FRI (prove/fri.rs) ----
  tm.begin(8);
  struct Round {
    uint32_t *evaluated, *nodes;
    size_t domain;
  };
  std::vector<Round> rounds;
  uint32_t* coeffs = final_coeffs;
  size_t size = n;
  while (size > B2Z_FRI_MIN_DEGREE) {
    Round r;
    r.domain = size * B2Z_INV_RATE;
    B2Z_TRY(mem.alloc(&r.evaluated, 4 * r.domain));
    B2Z_TRY(launch_evaluate(ctx, r.evaluated, coeffs, 4, ilog2(r.domain), 2));
    size_t rows = r.domain / B2Z_FRI_FOLD;
    B2Z_TRY(mem.alloc(&r.nodes, 2 * rows * 8));
    B2Z_TRY(merkle_build(ctx, r.nodes, r.evaluated, rows, B2Z_FRI_FOLD * 4));
    MerkleParams mp(rows, 64);
    B2Z_REQUIRE(seal.pos + 8 * mp.top_size <= seal.cap, "seal overflow");
    B2Z_CUDA(cudaMemcpyAsync(seal.w + seal.pos, r.nodes + 8 * mp.top_size, 32 * mp.top_size, cudaMemcpyDeviceToHost, ctx->stream));
    B2Z_CUDA(cudaMemcpyAsync(scratch, r.nodes + 8, 32, cudaMemcpyDeviceToHost, ctx->stream));
    B2Z_CUDA(b2z::stream_wait(ctx, ctx->stream));
    B2Z_TRY(check_remainders());
    seal.pos += 8 * mp.top_size;
    rng.mix(scratch);
    Fp4 fold_mix = rng.random_ext();
    uint32_t* out = nullptr;
    B2Z_TRY(mem.alloc(&out, 4 * size / B2Z_FRI_FOLD));
    B2Z_TRY(launch_fri_fold(ctx, out, coeffs, fold_mix, size / B2Z_FRI_FOLD));
    coeffs = out;
    size /= B2Z_FRI_FOLD;
    rounds.push_back(r);
  }
  {
    uint32_t* fin = nullptr;
    B2Z_TRY(mem.alloc(&fin, 4 * size));
    B2Z_CUDA(cudaMemcpyAsync(fin, coeffs, 16 * size, cudaMemcpyDeviceToDevice, ctx->stream));
    B2Z_TRY(launch_bit_reverse(ctx, fin, 4, ilog2(size)));
    B2Z_REQUIRE(seal.pos + 4 * size <= seal.cap, "seal overflow");
    B2Z_CUDA(cudaMemcpyAsync(seal.w + seal.pos, fin, 16 * size, cudaMemcpyDeviceToHost, ctx->stream));
    B2Z_CUDA(b2z::stream_wait(ctx, ctx->stream));
    B2Z_TRY(check_remainders());
    host_hash_elem_slice(seal.w + seal.pos, 4 * size, dg);
    seal.pos += 4 * size;
    rng.mix(dg);
  }
  tm.end(8);

  // ---- queries: positions from the transcript, one gather kernel for every tree ----
  tm.begin(9);
  {
    std::vector<uint32_t> positions(B2Z_QUERIES);
    for (auto& p : positions) p = rng.random_bits(ilog2(domain));
    std::vector<GatherTree> trees;
    size_t off = 0;
    auto add_tree = [&](const uint32_t* matrix, const uint32_t* nodes, size_t rows, size_t cols) {
      MerkleParams mp(rows, cols);
      GatherTree t{matrix, nodes, (uint32_t)rows, (uint32_t)cols, (uint32_t)mp.top_size, (uint32_t)off};
      trees.push_back(t);
      off += cols + mp.path_words();
    };
    for (int k = 0; k < 3; k++) {  // [RECALL med] opening order: protocol_params.h PP_QUERY_ORDER
      const int g = PP_QUERY_ORDER[k];
      add_tree(pg[g].evaluated, pg[g].nodes, domain, pg[g].count);
    }
    add_tree(check_pg.evaluated, check_pg.nodes, domain, B2Z_CHECK_SIZE);
    for (auto& r : rounds) add_tree(r.evaluated, r.nodes, r.domain / B2Z_FRI_FOLD, 64);
    size_t record_words = off;
    B2Z_REQUIRE(seal.pos + B2Z_QUERIES * record_words == seal.cap, "seal length mismatch");
    uint32_t *pos_dev = nullptr, *q_dev = nullptr;
    GatherTree* trees_dev = nullptr;
    B2Z_TRY(mem.alloc(&pos_dev, B2Z_QUERIES));
    B2Z_TRY(mem.alloc(&q_dev, B2Z_QUERIES * record_words));
    B2Z_TRY(mem.alloc(reinterpret_cast<uint32_t**>(&trees_dev), (sizeof(GatherTree) * trees.size() + 3) / 4));
    B2Z_CUDA(cudaMemcpyAsync(pos_dev, positions.data(), 4 * B2Z_QUERIES, cudaMemcpyHostToDevice, ctx->stream));
    B2Z_CUDA(cudaMemcpyAsync(trees_dev, trees.data(), sizeof(GatherTree) * trees.size(), cudaMemcpyHostToDevice, ctx->stream));
    B2Z_TRY(launch_gather_queries(ctx, q_dev, trees_dev, trees.size(), pos_dev, B2Z_QUERIES, record_words));
    B2Z_CUDA(cudaMemcpyAsync(seal.w + seal.pos, q_dev, 4 * B2Z_QUERIES * record_words, cudaMemcpyDeviceToHost, ctx->stream));
    seal.pos += B2Z_QUERIES * record_words;
  }
  tm.end(9);
  tm.end(10);
  B2Z_CUDA(b2z::stream_wait(ctx, ctx->stream));
  tm.collect();
  if (const char* tr = getenv("B2Z_TRACE")) {
    // one JSON line per segment on stderr (field names follow the reference harness: run-parallel.sh:15 `total_cycles`)
    if (tr[0] != '0') {
      fprintf(stderr, "{\"event\":\"segment_proved\",\"po2\":%u,\"total_cycles\":%zu,\"seal_words\":%zu", po2, n, seal_words);
      for (size_t i = 0; i < N_PHASES; i++) fprintf(stderr, ",\"%s_ms\":%.3f", PHASE_NAMES[i], ctx->phase_ms[i]);
      fprintf(stderr, "}\n");
    }
  }
  memcpy(seal_host, seal.w, 4 * seal_words);
  if (seal_words_out) *seal_words_out = seal_words;
  return nullptr;
}


// ---- host-buffer entry point: witness upload overlapped with the commits ----
// The witness crosses PCIe on ctx->copy_stream in column chunks, each followed by an event the compute stream waits on.
//   * pinned sources (cudaMallocHost / cudaHostRegister memory): the copies are truly asynchronous, the calling thread
//     enqueues all of them up front;
//   * pageable sources (a Rust Vec<u32>, a numpy array): cudaMemcpyAsync would stage them synchronously on the calling
//     thread and serialise the 3.5 GB upload with the compute, so STAGE_THREADS helper threads memcpy 16 MB pieces into a
//     pinned ring (2 slots per thread, owned by the context) and issue the device copies from there; the compute thread
//     blocks on a host-side latch until a chunk's event has been recorded.  Needs ~15 GB/s of host memcpy to stay hidden.
struct UploadGate {
  std::mutex m;
  std::condition_variable cv;
  std::vector<char> recorded;
  b2z_err error = nullptr;
  std::string error_text;
};
b2z_err upload_gate_wait(UploadGate* g, size_t idx) {
  std::unique_lock<std::mutex> lock(g->m);
  g->cv.wait(lock, [&] { return g->recorded[idx] || g->error; });
  return g->error ? set_error(g->error_text) : nullptr;
}

namespace {
constexpr size_t STAGE_THREADS = 4, STAGE_SLOTS = 2 * STAGE_THREADS, STAGE_PIECE_BYTES = (size_t)16 << 20;

struct UploadChunk {  // one event's worth of bytes
  char* dst;
  const char* src;
  size_t bytes;
  cudaEvent_t event;
  size_t first_piece = 0, n_pieces = 0;
  std::atomic<size_t> issued{0};
  UploadChunk(char* d, const char* s, size_t b, cudaEvent_t e) : dst(d), src(s), bytes(b), event(e) {}
  UploadChunk(const UploadChunk& o) : dst(o.dst), src(o.src), bytes(o.bytes), event(o.event), first_piece(o.first_piece), n_pieces(o.n_pieces) {}
};

bool host_pointer_is_pinned(const void* p) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

void staged_upload_worker(b2z_ctx* ctx, std::vector<UploadChunk>* chunks, std::atomic<size_t>* next_piece, size_t total_pieces, size_t tid,
                          UploadGate* gate) {
  cudaSetDevice(ctx->device);
  auto fail = [&](cudaError_t e) {
    std::lock_guard<std::mutex> lock(gate->m);
    if (!gate->error) {
      gate->error_text = std::string("staged witness upload: ") + cudaGetErrorString(e);
      gate->error = gate->error_text.c_str();
    }
    gate->cv.notify_all();
  };
  size_t use = 0;
  for (;;) {
    size_t p = next_piece->fetch_add(1);
    if (p >= total_pieces) return;
    // chunk of piece p (chunks are few: linear scan)
    size_t k = 0;
    while (k + 1 < chunks->size() && (*chunks)[k + 1].first_piece <= p) k++;
    UploadChunk& ch = (*chunks)[k];
    size_t off = (p - ch.first_piece) * STAGE_PIECE_BYTES;
    size_t len = ch.bytes - off < STAGE_PIECE_BYTES ? ch.bytes - off : STAGE_PIECE_BYTES;
    size_t slot = 2 * tid + (use++ & 1);
    cudaError_t e = cudaEventSynchronize(ctx->stage_events[slot]);  // the slot's previous device copy has drained
    if (e != cudaSuccess) return fail(e);
    char* ring = ctx->stage_ring + slot * STAGE_PIECE_BYTES;
    memcpy(ring, ch.src + off, len);
    e = cudaMemcpyAsync(ch.dst + off, ring, len, cudaMemcpyHostToDevice, ctx->copy_stream);
    if (e == cudaSuccess) e = cudaEventRecord(ctx->stage_events[slot], ctx->copy_stream);
    if (e != cudaSuccess) return fail(e);
    // the thread that issues a chunk's last piece records the chunk event: every earlier piece is already in the stream
    if (ch.issued.fetch_add(1) + 1 == ch.n_pieces) {
      e = cudaEventRecord(ch.event, ctx->copy_stream);
      if (e != cudaSuccess) return fail(e);
      std::lock_guard<std::mutex> lock(gate->m);
      gate->recorded[k] = 1;
      gate->cv.notify_all();
    }
  }
}
}  // namespace

// A witness on its way to the device (pinned sources): device buffers + the events its chunks are signalled on.  Begun early
// (the session prefetches segment i + 1 while segment i is being proved), consumed by prove_segment_uploaded.
struct HostUpload {
  uint32_t *code = nullptr, *data = nullptr, *noise = nullptr;
  WitnessReady ready;
  std::vector<cudaEvent_t> events;
};
static cudaEvent_t take_event(b2z_ctx* ctx) {
  if (!ctx->event_pool.empty()) {
    cudaEvent_t e = ctx->event_pool.back();
    ctx->event_pool.pop_back();
    return e;
  }
  cudaEvent_t e = nullptr;
  cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
  return e;
}
void host_upload_discard(b2z_ctx* ctx, HostUpload* up) {
  if (!up) return;
  cudaStreamSynchronize(ctx->copy_stream);
  for (void* p : {(void*)up->code, (void*)up->data, (void*)up->noise})
    if (p) cudaFreeAsync(p, ctx->stream);
  for (auto e : up->events) ctx->event_pool.push_back(e);
  delete up;
}
// upload plan in consumption order: code chunks (4 columns: the first compute starts after ~1 ms of PCIe), data chunks (32
// columns = two sponge blocks), accum noise
static void plan_chunks(const b2z_circuit* c, unsigned po2, size_t zk_rows, uint32_t* code, uint32_t* data, uint32_t* noise, const uint32_t* code_host,
                        const uint32_t* data_host, const uint32_t* noise_host, WitnessReady& ready, std::vector<UploadChunk>& chunks,
                        const std::function<cudaEvent_t()>& next_event) {
  const size_t n = (size_t)1 << po2;
  const size_t wa = c->group_size[0], wc = c->group_size[1], wd = c->group_size[2];
  // >= 256 MB of LDE per data chunk keeps the per-chunk launches (iNTT, LDE, persistent leaf-hash kernel) efficient
  ready.chunk_cols = po2 >= 21 ? 32 : 64;
  ready.code_chunk_cols = 4;
  const size_t nchunks = (wd + ready.chunk_cols - 1) / ready.chunk_cols, ncode = (wc + ready.code_chunk_cols - 1) / ready.code_chunk_cols;
  for (size_t k = 0; k < ncode; k++) {
    size_t c0 = k * ready.code_chunk_cols, cols = wc - c0 < ready.code_chunk_cols ? wc - c0 : ready.code_chunk_cols;
    cudaEvent_t e = next_event();
    chunks.emplace_back((char*)(code + c0 * n), (const char*)(code_host + c0 * n), 4 * cols * n, e);
    ready.code_chunks.push_back(e);
  }
  for (size_t k = 0; k < nchunks; k++) {
    size_t c0 = k * ready.chunk_cols, cols = wd - c0 < ready.chunk_cols ? wd - c0 : ready.chunk_cols;
    cudaEvent_t e = next_event();
    chunks.emplace_back((char*)(data + c0 * n), (const char*)(data_host + c0 * n), 4 * cols * n, e);
    ready.data_chunks.push_back(e);
  }
  ready.noise = next_event();
  chunks.emplace_back((char*)noise, (const char*)noise_host, 4 * wa * zk_rows, ready.noise);
  ready.code_gate0 = 0;
  ready.data_gate0 = ncode;
  ready.noise_gate = ncode + nchunks;
}
static bool sources_pinned(const b2z_circuit* c, unsigned po2, size_t zk_rows, const uint32_t* code_host, const uint32_t* data_host,
                           const uint32_t* noise_host) {
  (void)po2;
  const size_t wa = c->group_size[0], wc = c->group_size[1], wd = c->group_size[2];
  return (wc == 0 || host_pointer_is_pinned(code_host)) && (wd == 0 || host_pointer_is_pinned(data_host)) &&
         (wa * zk_rows == 0 || host_pointer_is_pinned(noise_host));
}

// Begin the upload of a PINNED witness: every copy and event is enqueued now; *out stays NULL for pageable sources.
b2z_err host_upload_begin(b2z_ctx* ctx, const b2z_circuit* c, unsigned po2, const uint32_t* code_host, const uint32_t* data_host,
                          const uint32_t* accum_noise_host, size_t zk_rows, HostUpload** out) {
  *out = nullptr;
  B2Z_REQUIRE(po2 >= 1 && po2 <= 24, "po2 out of range");
  const size_t n = (size_t)1 << po2;
  B2Z_REQUIRE(zk_rows < n, "zk_rows must be smaller than the trace");
  if (!sources_pinned(c, po2, zk_rows, code_host, data_host, accum_noise_host)) return nullptr;
  const size_t wa = c->group_size[0], wc = c->group_size[1], wd = c->group_size[2];
  if (!ctx->copy_stream) B2Z_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  HostUpload* up = new HostUpload();
  auto fail = [&](b2z_err e) {
    host_upload_discard(ctx, up);
    return e;
  };
  if (cudaMallocAsync(&up->code, 4 * (wc * n + 1), ctx->stream) != cudaSuccess || cudaMallocAsync(&up->data, 4 * (wd * n + 1), ctx->stream) != cudaSuccess ||
      cudaMallocAsync(&up->noise, 4 * (wa * zk_rows + 1), ctx->stream) != cudaSuccess)
    return fail(set_error(std::string("witness buffers: ") + cudaGetErrorString(cudaGetLastError())));
  // the copy stream may only touch the buffers once the (stream-ordered) allocations are visible
  cudaEvent_t alloc_done = take_event(ctx);
  up->events.push_back(alloc_done);
  cudaEventRecord(alloc_done, ctx->stream);
  cudaStreamWaitEvent(ctx->copy_stream, alloc_done, 0);
  std::vector<UploadChunk> chunks;
  plan_chunks(c, po2, zk_rows, up->code, up->data, up->noise, code_host, data_host, accum_noise_host, up->ready, chunks, [&] {
    cudaEvent_t e = take_event(ctx);
    up->events.push_back(e);
    return e;
  });
  for (auto& ch : chunks) {
    if (ch.bytes && cudaMemcpyAsync(ch.dst, ch.src, ch.bytes, cudaMemcpyHostToDevice, ctx->copy_stream) != cudaSuccess)
      return fail(set_error(std::string("witness upload: ") + cudaGetErrorString(cudaGetLastError())));
    cudaEventRecord(ch.event, ctx->copy_stream);
  }
  *out = up;
  return nullptr;
}

b2z_err prove_segment_uploaded(b2z_ctx* ctx, const b2z_circuit* c, unsigned po2, HostUpload* up, const uint32_t* globals_host, size_t zk_rows,
                               uint32_t* seal_host, size_t* seal_words) {
  ctx->last_upload_staged = false;
  b2z_err err = prove_segment_device(ctx, c, po2, up->code, up->data, globals_host, up->noise, zk_rows, seal_host, seal_words, &up->ready);
  host_upload_discard(ctx, up);
  return err;
}

b2z_err prove_segment_host(b2z_ctx* ctx, const b2z_circuit* c, unsigned po2, const uint32_t* code_host, const uint32_t* data_host,
                           const uint32_t* globals_host, const uint32_t* accum_noise_host, size_t zk_rows, uint32_t* seal_host,
                           size_t* seal_words) {
  HostUpload* up = nullptr;
  B2Z_TRY(host_upload_begin(ctx, c, po2, code_host, data_host, accum_noise_host, zk_rows, &up));
  if (up) return prove_segment_uploaded(ctx, c, po2, up, globals_host, zk_rows, seal_host, seal_words);
  // pageable sources: staged through the pinned ring by helper threads
  const size_t n = (size_t)1 << po2;
  const size_t wa = c->group_size[0], wc = c->group_size[1], wd = c->group_size[2];
  if (!ctx->copy_stream) B2Z_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
  ctx->last_upload_staged = true;
  DevMem wit(ctx);  // released on every exit path
  uint32_t *code = nullptr, *data = nullptr, *noise = nullptr;
  B2Z_TRY(wit.alloc(&code, wc * n + 1));
  B2Z_TRY(wit.alloc(&data, wd * n + 1));
  B2Z_TRY(wit.alloc(&noise, wa * zk_rows + 1));
  std::vector<cudaEvent_t> evs;
  struct Return {
    b2z_ctx* ctx;
    std::vector<cudaEvent_t>& evs;
    ~Return() {
      for (auto e : evs) ctx->event_pool.push_back(e);
    }
  } ret{ctx, evs};
  cudaEvent_t alloc_done = take_event(ctx);
  evs.push_back(alloc_done);
  B2Z_CUDA(cudaEventRecord(alloc_done, ctx->stream));
  B2Z_CUDA(cudaStreamWaitEvent(ctx->copy_stream, alloc_done, 0));
  WitnessReady ready;
  std::vector<UploadChunk> chunks;
  plan_chunks(c, po2, zk_rows, code, data, noise, code_host, data_host, accum_noise_host, ready, chunks, [&] {
    cudaEvent_t e = take_event(ctx);
    evs.push_back(e);
    return e;
  });
  if (!ctx->stage_ring) {
    B2Z_CUDA(cudaMallocHost((void**)&ctx->stage_ring, STAGE_SLOTS * STAGE_PIECE_BYTES));
    for (size_t i = 0; i < STAGE_SLOTS; i++) {
      cudaEvent_t e;
      B2Z_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      ctx->stage_events.push_back(e);
    }
  }
  size_t total_pieces = 0;
  for (auto& ch : chunks) {
    ch.first_piece = total_pieces;
    ch.n_pieces = (ch.bytes + STAGE_PIECE_BYTES - 1) / STAGE_PIECE_BYTES;
    if (ch.n_pieces == 0) ch.n_pieces = 1;  // an empty chunk still records its event (zero-byte piece)
    total_pieces += ch.n_pieces;
  }
  UploadGate gate;
  gate.recorded.assign(chunks.size(), 0);
  ready.gate = &gate;
  std::atomic<size_t> next_piece{0};
  std::vector<std::thread> workers;
  for (size_t t = 0; t < STAGE_THREADS; t++) workers.emplace_back(staged_upload_worker, ctx, &chunks, &next_piece, total_pieces, t, &gate);
  b2z_err err = prove_segment_device(ctx, c, po2, code, data, globals_host, noise, zk_rows, seal_host, seal_words, &ready);
  if (err) next_piece.store(total_pieces);  // stop issuing pieces; the workers drain and exit
  for (auto& w : workers) w.join();
  cudaStreamSynchronize(ctx->copy_stream);
  return err;
}

// Control root of a code group (risc0: the control ID of a po2): the Merkle root its commitment has in every honest seal.
// Same commit sequence as the prover's code group (iNTT, zk_shift, x4 LDE, Poseidon2 tree); the root is what
// b2z_verify_segment compares the seal's code root with.
b2z_err compute_control_root(b2z_ctx* ctx, unsigned po2, const uint32_t* code_dev, size_t cols, uint32_t* root_host) {
  B2Z_REQUIRE(po2 >= 1 && po2 <= 24 && cols > 0, "control root: bad arguments");
  const size_t n = (size_t)1 << po2, domain = 4 * n;
  DevMem mem(ctx);
  uint32_t *coeffs = nullptr, *evaluated = nullptr, *nodes = nullptr;
  B2Z_TRY(mem.alloc(&coeffs, cols * n));
  B2Z_TRY(mem.alloc(&evaluated, cols * domain));
  B2Z_TRY(mem.alloc(&nodes, 2 * domain * 8));
  B2Z_TRY(launch_interpolate(ctx, coeffs, cols, po2, /*fuse_zk_shift=*/true, code_dev));
  B2Z_TRY(launch_evaluate(ctx, evaluated, coeffs, cols, po2 + 2, 2));
  B2Z_TRY(merkle_build(ctx, nodes, evaluated, domain, cols));
  B2Z_CUDA(cudaMemcpyAsync(root_host, nodes + 8, 32, cudaMemcpyDeviceToHost, ctx->stream));
  B2Z_CUDA(b2z::stream_wait(ctx, ctx->stream));
  return nullptr;
}

}  // namespace b2z
