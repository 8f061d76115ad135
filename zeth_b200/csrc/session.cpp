// Session level of the prover: what zeth reaches through `default_prover().prove(env, elf)`
// (/root/reference crates/host/src/lib.rs:131-142) and checks with `receipt.verify(image_id)` + `receipt.journal`
// (crates/host/src/bin/cli.rs:103-107).
//
// Upstream (risc0-zkvm 3.0.3 host/server/prove/prover_impl.rs) runs `prove_session`: a sequential loop
// `for segment in session.segments { prove_segment(segment) -> SegmentReceipt }` and wraps the list as
// Receipt{inner: Composite{segments, ..}, journal}.  Segments are independent, so here a session owns `in_flight` workers
// on one GPU (each its own context = stream + tables + circuit-kernel instance) that pull segments from a queue: while one
// worker waits for a 32-byte transcript round trip, the other's kernels fill the GPU.  Receipts come back in index order.
//
// What this file can and cannot reproduce of upstream's containers: the SEAL inside every SegmentReceipt is the bit-exact
// part (Vec<u32> transcript).  The container encoding (upstream: serde/bincode of the Rust structs) and the ReceiptClaim
// (SHA-256 tagged structs over the rv32im globals) are not recallable byte-for-byte offline; the container here is a
// documented little-endian framing of the same fields, and the claim is the pre/post state digest pair that the synthetic
// circuits carry in their globals (include/protocol_params.h "synthetic sessions").
#include <atomic>
#include <chrono>
#include <cstring>
#include <mutex>
#include <thread>

#include "circuit.h"
#include "common.h"

namespace b2z {
b2z_err prove_segment_device(b2z_ctx* ctx, const b2z_circuit* c, unsigned po2, const uint32_t* code_dev, const uint32_t* data_dev,
                             const uint32_t* globals_host, const uint32_t* accum_noise_dev, size_t zk_rows, uint32_t* seal_host,
                             size_t* seal_words_out, const struct WitnessReady* ready);
b2z_err prove_segment_host(b2z_ctx* ctx, const b2z_circuit* c, unsigned po2, const uint32_t* code_host, const uint32_t* data_host,
                           const uint32_t* globals_host, const uint32_t* accum_noise_host, size_t zk_rows, uint32_t* seal_host,
                           size_t* seal_words);
int verify_segment(const b2z_circuit* c, const uint32_t* seal, size_t seal_words);
struct HostUpload;
b2z_err host_upload_begin(b2z_ctx* ctx, const b2z_circuit* c, unsigned po2, const uint32_t* code_host, const uint32_t* data_host,
                          const uint32_t* accum_noise_host, size_t zk_rows, HostUpload** out);
b2z_err prove_segment_uploaded(b2z_ctx* ctx, const b2z_circuit* c, unsigned po2, HostUpload* up, const uint32_t* globals_host, size_t zk_rows,
                               uint32_t* seal_host, size_t* seal_words);
void host_upload_discard(b2z_ctx* ctx, HostUpload* up);

// digest naming (proof system, circuit): SegmentReceipt::verifier_parameters stand-in
void verifier_parameters(const b2z_circuit* c, uint32_t* out8) {
  uint32_t e[16], a[8], b[8];
  for (int i = 0; i < 16; i++) e[i] = fp_from_u32((uint8_t)PP_PROOF_SYSTEM_INFO[i]);
  host_hash_elem_slice(e, 16, a);
  for (int i = 0; i < 16; i++) e[i] = fp_from_u32(c->info[i]);
  host_hash_elem_slice(e, 16, b);
  host_hash_pair(a, b, out8);
}

// claim words of a seal: the header (globals || po2) opens the seal; pre / post digests sit in the last 16 globals
static bool decode_claim(const b2z_circuit* c, const uint32_t* seal, size_t seal_words, uint32_t* pre8, uint32_t* post8, uint32_t* po2) {
  const size_t g = c->globals_size;
  if (seal_words < g + 1) return false;
  *po2 = fp_to_u32(seal[g]);
  memset(pre8, 0, 32);
  memset(post8, 0, 32);
  if (pp_claim_pre_offset(g) >= 0) {
    memcpy(pre8, seal + pp_claim_pre_offset(g), 32);
    memcpy(post8, seal + pp_claim_post_offset(g), 32);
  }
  return true;
}

static void free_receipt_contents(b2z_receipt* r) {
  if (!r) return;
  if (r->segments) {
    for (size_t i = 0; i < r->n_segments; i++) free(r->segments[i].seal);
    free(r->segments);
  }
  free(r->journal);
}

b2z_err receipt_assemble(const b2z_circuit* c, const uint32_t* const* seals, const size_t* seal_words, size_t n, const uint8_t* journal,
                         size_t journal_len, b2z_receipt** out) {
  B2Z_REQUIRE(c && out && (n == 0 || (seals && seal_words)), "receipt_assemble: null argument");
  b2z_receipt* r = (b2z_receipt*)calloc(1, sizeof(b2z_receipt));
  B2Z_REQUIRE(r, "out of memory");
  r->segments = (b2z_segment_receipt*)calloc(n ? n : 1, sizeof(b2z_segment_receipt));
  r->n_segments = r->segments ? n : 0;
  verifier_parameters(c, r->verifier_parameters);
  if (journal_len) r->journal = (uint8_t*)malloc(journal_len);
  if (!r->segments || (journal_len && (!r->journal || !journal))) {
    free_receipt_contents(r);
    free(r);
    return set_error(journal_len && !journal ? "receipt_assemble: null journal" : "out of memory");
  }
  if (journal_len) {
    memcpy(r->journal, journal, journal_len);
    r->journal_len = journal_len;
  }
  for (size_t i = 0; i < n; i++) {
    b2z_segment_receipt& s = r->segments[i];
    s.index = (uint32_t)i;
    snprintf(s.hashfn, sizeof s.hashfn, "poseidon2");
    memcpy(s.verifier_parameters, r->verifier_parameters, 32);
    if (!decode_claim(c, seals[i], seal_words[i], s.claim_pre, s.claim_post, &s.po2)) {
      free_receipt_contents(r);
      free(r);
      return set_error("receipt_assemble: seal " + std::to_string(i) + " is shorter than its header");
    }
    s.seal = (uint32_t*)malloc(4 * seal_words[i]);
    if (!s.seal) {
      free_receipt_contents(r);
      free(r);
      return set_error("out of memory");
    }
    memcpy(s.seal, seals[i], 4 * seal_words[i]);
    s.seal_words = seal_words[i];
  }
  *out = r;
  return nullptr;
}

// receipt.verify(image_id): every SegmentReceipt in index order, seal verified against the circuit (incl. control root),
// claim == what the seal's own header says, claims chained pre -> post, first pre-state == image id.
int verify_receipt(const b2z_circuit* c, const b2z_receipt* r, const uint32_t* image_id, size_t* bad_segment) {
  if (bad_segment) *bad_segment = 0;
  if (!c || !r || !r->segments || r->n_segments == 0) return 101;
  uint32_t vp[8];
  verifier_parameters(c, vp);
  if (memcmp(vp, r->verifier_parameters, 32) != 0) return 104;
  const bool has_claim = pp_claim_pre_offset(c->globals_size) >= 0;
  for (size_t i = 0; i < r->n_segments; i++) {
    const b2z_segment_receipt& s = r->segments[i];
    if (bad_segment) *bad_segment = i;
    if (s.index != i) return 102;
    if (strncmp(s.hashfn, "poseidon2", sizeof s.hashfn) != 0) return 103;
    if (memcmp(vp, s.verifier_parameters, 32) != 0) return 104;
    if (!s.seal) return 101;
    int rc = verify_segment(c, s.seal, s.seal_words);
    if (rc != 0) return rc;
    uint32_t pre[8], post[8], po2 = 0;
    if (!decode_claim(c, s.seal, s.seal_words, pre, post, &po2)) return 105;
    if (po2 != s.po2 || memcmp(pre, s.claim_pre, 32) != 0 || memcmp(post, s.claim_post, 32) != 0) return 105;
    if (has_claim) {
      if (i == 0) {
        if (image_id && memcmp(image_id, s.claim_pre, 32) != 0) return 107;
      } else if (memcmp(r->segments[i - 1].claim_post, s.claim_pre, 32) != 0) {
        return 106;
      }
    }
  }
  if (bad_segment) *bad_segment = 0;
  return 0;
}

// ---- container (de)serialisation: little-endian framing, see the header comment ----
static const uint32_t RECEIPT_MAGIC = 0x525a3242u;  // "B2ZR"
struct Writer {
  std::vector<uint8_t> b;
  void u32(uint32_t v) {
    for (int i = 0; i < 4; i++) b.push_back((uint8_t)(v >> (8 * i)));
  }
  void u64(uint64_t v) {
    u32((uint32_t)v);
    u32((uint32_t)(v >> 32));
  }
  void words(const uint32_t* w, size_t n) {
    for (size_t i = 0; i < n; i++) u32(w[i]);
  }
  void bytes(const void* p, size_t n) {
    const uint8_t* q = (const uint8_t*)p;
    b.insert(b.end(), q, q + n);
  }
};
struct Reader {
  const uint8_t* p;
  size_t len, pos = 0;
  bool ok = true;
  bool need(size_t n) {
    if (!ok || len - pos < n) ok = false;
    return ok;
  }
  uint32_t u32() {
    if (!need(4)) return 0;
    uint32_t v = 0;
    for (int i = 0; i < 4; i++) v |= (uint32_t)p[pos + i] << (8 * i);
    pos += 4;
    return v;
  }
  uint64_t u64() {
    uint64_t lo = u32();
    return lo | ((uint64_t)u32() << 32);
  }
  void words(uint32_t* w, size_t n) {
    for (size_t i = 0; i < n; i++) w[i] = u32();
  }
  void bytes(void* dst, size_t n) {
    if (!need(n)) return;
    memcpy(dst, p + pos, n);
    pos += n;
  }
};

b2z_err receipt_serialize(const b2z_receipt* r, uint8_t** bytes, size_t* len) {
  B2Z_REQUIRE(r && bytes && len, "receipt_serialize: null argument");
  Writer w;
  w.u32(RECEIPT_MAGIC);
  w.u32(1);  // version
  w.words(r->verifier_parameters, 8);
  w.u64(r->n_segments);
  w.u64(r->journal_len);
  w.bytes(r->journal, r->journal_len);
  for (size_t i = 0; i < r->n_segments; i++) {
    const b2z_segment_receipt& s = r->segments[i];
    w.u32(s.index);
    w.u32(s.po2);
    w.bytes(s.hashfn, sizeof s.hashfn);
    w.words(s.verifier_parameters, 8);
    w.words(s.claim_pre, 8);
    w.words(s.claim_post, 8);
    w.u64(s.seal_words);
    w.words(s.seal, s.seal_words);
  }
  *bytes = (uint8_t*)malloc(w.b.size() ? w.b.size() : 1);
  B2Z_REQUIRE(*bytes, "out of memory");
  memcpy(*bytes, w.b.data(), w.b.size());
  *len = w.b.size();
  return nullptr;
}

b2z_err receipt_deserialize(const uint8_t* bytes, size_t len, b2z_receipt** out) {
  B2Z_REQUIRE(bytes && out, "receipt_deserialize: null argument");
  Reader rd{bytes, len};
  B2Z_REQUIRE(rd.u32() == RECEIPT_MAGIC && rd.u32() == 1 && rd.ok, "receipt: bad magic / version");
  b2z_receipt* r = (b2z_receipt*)calloc(1, sizeof(b2z_receipt));
  B2Z_REQUIRE(r, "out of memory");
  auto fail = [&](const char* msg) {
    free_receipt_contents(r);
    free(r);
    return set_error(msg);
  };
  rd.words(r->verifier_parameters, 8);
  uint64_t nseg = rd.u64(), jlen = rd.u64();
  // every segment needs at least its fixed fields: bounds the allocation by the input length
  if (!rd.ok || jlen > len - rd.pos || nseg > (len - rd.pos) / 128 + 1) return fail("receipt: truncated header");
  if (jlen) {
    r->journal = (uint8_t*)malloc(jlen);
    if (!r->journal) return fail("out of memory");
    rd.bytes(r->journal, jlen);
    r->journal_len = jlen;
  }
  r->segments = (b2z_segment_receipt*)calloc(nseg ? nseg : 1, sizeof(b2z_segment_receipt));
  if (!r->segments) return fail("out of memory");
  for (uint64_t i = 0; i < nseg; i++) {
    b2z_segment_receipt& s = r->segments[i];
    r->n_segments = i + 1;
    s.index = rd.u32();
    s.po2 = rd.u32();
    rd.bytes(s.hashfn, sizeof s.hashfn);
    s.hashfn[sizeof s.hashfn - 1] = 0;
    rd.words(s.verifier_parameters, 8);
    rd.words(s.claim_pre, 8);
    rd.words(s.claim_post, 8);
    uint64_t sw = rd.u64();
    if (!rd.ok || sw > (len - rd.pos) / 4) return fail("receipt: truncated segment");
    s.seal = (uint32_t*)malloc(sw ? 4 * sw : 4);
    if (!s.seal) return fail("out of memory");
    rd.words(s.seal, sw);
    s.seal_words = sw;
  }
  if (!rd.ok || rd.pos != len) return fail("receipt: trailing or missing bytes");
  *out = r;
  return nullptr;
}

}  // namespace b2z

// ---- the session object ----
struct b2z_session {
  int device = 0;
  struct Worker {
    b2z_ctx* ctx = nullptr;
    b2z_circuit* circuit = nullptr;
  };
  std::vector<Worker> workers;
  double last_wall_ms = 0;
  size_t last_retries = 0, last_cache_hits = 0;
  std::string seal_cache_dir;  // optional on-disk seal cache (resume of a long session)
  std::vector<double> last_segment_ms;
};

using namespace b2z;

extern "C" {

b2z_err b2z_session_create(int device, const uint32_t* blob, size_t words, unsigned in_flight, b2z_session** out) {
  B2Z_REQUIRE(blob && out, "null argument");
  B2Z_REQUIRE(in_flight >= 1 && in_flight <= 8, "in_flight must be 1..8");
  b2z_session* s = new b2z_session();
  s->device = device;
  for (unsigned i = 0; i < in_flight; i++) {
    b2z_session::Worker w;
    b2z_err e = b2z_init(device, &w.ctx);
    // every worker loads its own instance of the circuit kernels: the generated eval_check keeps its tables in module
    // __constant__ memory, which two segments in flight must not share
    if (!e) e = b2z_circuit_load(w.ctx, blob, words, &w.circuit);
    if (e) {
      std::string msg = e;
      if (w.ctx) b2z_destroy(w.ctx);
      b2z_session_destroy(s);
      return set_error(msg);
    }
    s->workers.push_back(w);
  }
  *out = s;
  return nullptr;
}

b2z_err b2z_session_destroy(b2z_session* s) {
  if (!s) return nullptr;
  for (auto& w : s->workers) {
    if (w.circuit) b2z_circuit_free(w.circuit);
    if (w.ctx) b2z_destroy(w.ctx);
  }
  delete s;
  return nullptr;
}

b2z_circuit* b2z_session_circuit(b2z_session* s) { return (s && !s->workers.empty()) ? s->workers[0].circuit : nullptr; }
b2z_ctx* b2z_session_context(b2z_session* s, unsigned worker) { return (s && worker < s->workers.size()) ? s->workers[worker].ctx : nullptr; }
unsigned b2z_session_in_flight(const b2z_session* s) { return s ? (unsigned)s->workers.size() : 0; }
uint64_t b2z_session_launch_count(const b2z_session* s) {
  uint64_t n = 0;
  if (s)
    for (auto& w : s->workers) n += w.ctx->launches;
  return n;
}
double b2z_session_last_wall_ms(const b2z_session* s) { return s ? s->last_wall_ms : 0.0; }
size_t b2z_session_last_retries(const b2z_session* s) { return s ? s->last_retries : 0; }
size_t b2z_session_last_segment_ms(const b2z_session* s, double* ms, size_t cap) {
  if (!s) return 0;
  for (size_t i = 0; i < s->last_segment_ms.size() && i < cap; i++) ms[i] = s->last_segment_ms[i];
  return s->last_segment_ms.size();
}

// Seal cache (SURVEY.md 5 "checkpoint / resume"; the reference's analogue is its input cache, crates/host/src/bin/cli.rs:113-145):
// with a directory set, every proved segment's seal is written to <dir>/seal_<circuit>_<po2>_<index>_<header digest>.bin and a
// later session that meets the same (circuit, po2, index, header = globals || po2) loads it instead of proving -- a 500-segment
// job that died at segment 400 resumes there.  The header digest ties the file to the segment's public state (pre / post
// digests live in the globals); a cached seal is re-verified before it is trusted when a control root is registered.
b2z_err b2z_session_set_seal_cache(b2z_session* s, const char* dir) {
  B2Z_REQUIRE(s, "null session");
  s->seal_cache_dir = dir ? dir : "";
  return nullptr;
}
size_t b2z_session_last_cache_hits(const b2z_session* s) { return s ? s->last_cache_hits : 0; }

static std::string seal_cache_path(const b2z_session* s, const b2z_circuit* c, const b2z_segment_input& g) {
  uint32_t vp[8], hd[8];
  verifier_parameters(c, vp);
  std::vector<uint32_t> header(g.globals, g.globals + c->globals_size);
  header.push_back(fp_from_u32(g.po2));
  host_hash_elem_slice(header.data(), header.size(), hd);
  char name[160];
  snprintf(name, sizeof name, "/seal_%08x%08x_%u_%u_%08x%08x%08x%08x.bin", vp[0], vp[1], g.po2, g.index, hd[0], hd[1], hd[2], hd[3]);
  return s->seal_cache_dir + name;
}
static bool seal_cache_load(const std::string& path, std::vector<uint32_t>& seal) {
  FILE* f = fopen(path.c_str(), "rb");
  if (!f) return false;
  bool ok = fread(seal.data(), 4, seal.size(), f) == seal.size() && fgetc(f) == EOF;
  fclose(f);
  return ok;
}
static void seal_cache_store(const std::string& path, const std::vector<uint32_t>& seal) {
  std::string tmp = path + ".tmp." + std::to_string((unsigned long long)(uintptr_t)&seal);
  FILE* f = fopen(tmp.c_str(), "wb");
  if (!f) return;
  bool ok = fwrite(seal.data(), 4, seal.size(), f) == seal.size();
  ok = fclose(f) == 0 && ok;
  if (ok) ok = rename(tmp.c_str(), path.c_str()) == 0;
  if (!ok) remove(tmp.c_str());
}

b2z_err b2z_session_set_control_root(b2z_session* s, unsigned po2, const uint32_t* root8) {
  B2Z_REQUIRE(s && root8, "null argument");
  for (auto& w : s->workers) B2Z_TRY(b2z_circuit_set_control_root(w.circuit, po2, root8));
  return nullptr;
}

// prove_session: segments in, Receipt (ordered SegmentReceipts) out
b2z_err b2z_prove_session(b2z_session* s, const b2z_segment_input* segs, size_t n, const uint8_t* journal, size_t journal_len,
                          b2z_receipt** out) {
  B2Z_REQUIRE(s && out && (n == 0 || segs), "null argument");
  B2Z_REQUIRE(!s->workers.empty(), "session has no workers");
  const b2z_circuit* c0 = s->workers[0].circuit;
  std::vector<std::vector<uint32_t>> seals(n);
  for (size_t i = 0; i < n; i++) {
    B2Z_REQUIRE(segs[i].po2 >= 1 && segs[i].po2 <= 24, "segment po2 out of range");
    seals[i].resize(circuit_seal_words(c0, segs[i].po2));
  }
  s->last_segment_ms.assign(n, 0.0);
  std::atomic<size_t> next{0}, retries{0}, cache_hits{0};
  // seals already on disk are taken first (single-threaded, host only): the workers then see them as done
  std::vector<char> done(n, 0);
  if (!s->seal_cache_dir.empty())
    for (size_t i = 0; i < n; i++) {
      if (!seal_cache_load(seal_cache_path(s, c0, segs[i]), seals[i])) continue;
      const bool bound = segs[i].po2 < 32 && c0->has_control_root[segs[i].po2];
      if (bound && verify_segment(c0, seals[i].data(), seals[i].size()) != 0) continue;  // stale / corrupt file: prove again
      done[i] = 1;
      cache_hits.fetch_add(1);
    }
  std::mutex err_mutex;
  std::string first_error;
  std::atomic<bool> failed{false};
  auto t0 = std::chrono::steady_clock::now();
  auto work = [&](size_t wi) {
    b2z_session::Worker& w = s->workers[wi];
    cudaSetDevice(s->device);
    // the worker claims its NEXT segment early and starts that witness's upload (pinned host sources) before it proves the
    // current one: PCIe runs decoupled from the commit phases, and two workers do not fall into upload-bound lockstep
    auto claim = [&]() {  // next segment that still has to be proved
      size_t i;
      do i = next.fetch_add(1);
      while (i < n && done[i]);
      return i;
    };
    auto begin = [&](size_t i, HostUpload** up) {
      *up = nullptr;
      if (i < n && !segs[i].on_device && !failed.load())
        host_upload_begin(w.ctx, w.circuit, segs[i].po2, segs[i].code, segs[i].data, segs[i].accum_noise, segs[i].zk_rows, up);  // NULL on pageable sources / errors
    };
    size_t i_next = claim();
    HostUpload* up_next = nullptr;
    begin(i_next, &up_next);
    for (;;) {
      const size_t i = i_next;
      HostUpload* up = up_next;
      if (i >= n || failed.load()) {
        host_upload_discard(w.ctx, up);
        return;
      }
      i_next = claim();
      begin(i_next, &up_next);
      const b2z_segment_input& g = segs[i];
      b2z_err e = nullptr;
      auto ts = std::chrono::steady_clock::now();
      for (int attempt = 0; attempt < 2; attempt++) {  // one retry: a transient device error must not lose the session
        size_t sw = 0;
        if (g.on_device) {
          e = prove_segment_device(w.ctx, w.circuit, g.po2, g.code, g.data, g.globals, g.accum_noise, g.zk_rows, seals[i].data(), &sw, nullptr);
        } else if (up) {
          e = prove_segment_uploaded(w.ctx, w.circuit, g.po2, up, g.globals, g.zk_rows, seals[i].data(), &sw);
          up = nullptr;
        } else {
          e = prove_segment_host(w.ctx, w.circuit, g.po2, g.code, g.data, g.globals, g.accum_noise, g.zk_rows, seals[i].data(), &sw);
        }
        if (!e) break;
        // an unsatisfied witness is deterministic: retrying cannot help
        if (strstr(e, "does not satisfy")) break;
        if (attempt == 0) {
          retries.fetch_add(1);
          b2z::stream_wait(w.ctx, w.ctx->stream);
          cudaGetLastError();
        }
      }
      s->last_segment_ms[i] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - ts).count();
      if (!e && !s->seal_cache_dir.empty()) seal_cache_store(seal_cache_path(s, w.circuit, g), seals[i]);
      if (e) {
        std::lock_guard<std::mutex> lock(err_mutex);
        if (first_error.empty()) first_error = "segment " + std::to_string(g.index) + ": " + e;
        failed.store(true);
        host_upload_discard(w.ctx, up_next);
        return;
      }
    }
  };
  size_t nthreads = s->workers.size() < n ? s->workers.size() : n;
  if (nthreads <= 1) {
    if (n) work(0);
  } else {
    std::vector<std::thread> threads;
    for (size_t wi = 0; wi < nthreads; wi++) threads.emplace_back(work, wi);
    for (auto& t : threads) t.join();
  }
  s->last_wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
  s->last_retries = retries.load();
  s->last_cache_hits = cache_hits.load();
  if (failed.load()) return set_error(first_error);
  std::vector<const uint32_t*> ptrs(n);
  std::vector<size_t> lens(n);
  for (size_t i = 0; i < n; i++) {
    ptrs[i] = seals[i].data();
    lens[i] = seals[i].size();
  }
  B2Z_TRY(receipt_assemble(c0, ptrs.data(), lens.data(), n, journal, journal_len, out));
  for (size_t i = 0; i < n; i++) (*out)->segments[i].index = segs[i].index;
  return nullptr;
}

b2z_err b2z_receipt_assemble(const b2z_circuit* c, const uint32_t* const* seals, const size_t* seal_words, size_t n, const uint8_t* journal,
                             size_t journal_len, b2z_receipt** out) {
  return receipt_assemble(c, seals, seal_words, n, journal, journal_len, out);
}
b2z_err b2z_receipt_free(b2z_receipt* r) {
  if (!r) return nullptr;
  free_receipt_contents(r);
  free(r);
  return nullptr;
}
b2z_err b2z_receipt_serialize(const b2z_receipt* r, uint8_t** bytes, size_t* len) { return receipt_serialize(r, bytes, len); }
b2z_err b2z_receipt_deserialize(const uint8_t* bytes, size_t len, b2z_receipt** out) { return receipt_deserialize(bytes, len, out); }
void b2z_bytes_free(uint8_t* bytes) { free(bytes); }
int b2z_verify_receipt(const b2z_circuit* c, const b2z_receipt* r, const uint32_t* image_id, size_t* bad_segment) {
  return verify_receipt(c, r, image_id, bad_segment);
}

}  // extern "C"
