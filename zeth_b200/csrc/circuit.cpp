// Circuit blob parsing, TapSet derivation, CUDA code generation for the constraint program and NVRTC loading.
//
// Shape follows risc0-zkp 3.0.2 taps.rs (TapSet: taps sorted by (group, offset, back); registers; combos) and
// adapter.rs (PolyExtStep program) -- SURVEY.md Appendix A.7.  Upstream ships the rv32im program as *generated*
// CUDA (risc0-circuit-rv32im-sys kernels/cuda/eval_check*.cu).  Here the generator is part of the library: the
// program arrives as data (circuit blob), is specialised into one straight-line sm_100a kernel by NVRTC (cached on
// disk as a cubin keyed by the source hash) and loaded with cudaLibraryLoadData.
#include "circuit.h"

#include <dlfcn.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstring>
#include <fstream>
#include <sstream>

namespace b2z {

static thread_local std::string g_err;
const char* set_error(const std::string& msg) {
  g_err = msg;
  return g_err.c_str();
}

static int cmp_backs(const uint32_t* a, size_t na, const uint32_t* b, size_t nb) {
  size_t n = na < nb ? na : nb;
  for (size_t i = 0; i < n; i++)
    if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
  return na == nb ? 0 : (na < nb ? -1 : 1);
}

b2z_err circuit_parse(b2z_circuit* c, const uint32_t* blob, size_t words) {
  B2Z_REQUIRE(words >= 24 && blob[0] == 0x435a3242u && blob[1] == 1, "circuit blob: bad magic/version");
  for (int g = 0; g < 3; g++) c->group_size[g] = blob[2 + g];
  c->globals_size = blob[5];
  c->mix_size = blob[6];
  c->n_taps = blob[7];
  size_t n_steps = blob[8];
  c->ret = blob[9];
  memcpy(c->info, blob + 10, 16);
  size_t n_wit_steps = blob[14], n_wit_stores = blob[15], n_accum = blob[16], n_pins = blob[17];
  size_t pos = 24;
  size_t need = pos + c->n_taps + 5 * n_steps + 5 * n_wit_steps + 2 * n_wit_stores + n_accum + 2 * n_pins + c->group_size[2];
  B2Z_REQUIRE(words >= need, "circuit blob: truncated");
  for (size_t i = 0; i < c->n_taps; i++) {
    uint32_t w = blob[pos++];
    c->tap_offset.push_back(w & 0xffff);
    c->tap_back.push_back((w >> 16) & 0xff);
    c->tap_group.push_back(w >> 24);
    B2Z_REQUIRE((w >> 24) < 3 && (w & 0xffff) < c->group_size[w >> 24], "circuit blob: tap out of range");
  }
  auto read_steps = [&](std::vector<b2z_circuit::Step>& v, size_t n) {
    for (size_t i = 0; i < n; i++) {
      v.push_back({blob[pos], blob[pos + 1], blob[pos + 2], blob[pos + 3], blob[pos + 4]});
      pos += 5;
    }
  };
  read_steps(c->steps, n_steps);
  read_steps(c->wit_steps, n_wit_steps);
  for (size_t i = 0; i < n_wit_stores; i++) {
    c->wit_store_col.push_back(blob[pos++]);
    c->wit_store_var.push_back(blob[pos++]);
  }
  for (size_t i = 0; i < n_accum; i++) c->accum_col.push_back(blob[pos++]);
  for (size_t i = 0; i < n_pins; i++) {
    c->pin_col.push_back(blob[pos++]);
    c->pin_global.push_back(blob[pos++]);
  }
  for (size_t i = 0; i < c->group_size[2]; i++) {
    uint32_t k = blob[pos++];
    B2Z_REQUIRE(k <= 2 || ((k & 0xff) == 3 && ((k >> 16) & 0xff) > ((k >> 8) & 0xff)), "circuit blob: bad data column kind");
    c->col_kind.push_back(k);
  }
  B2Z_REQUIRE(c->mix_size >= 4 * n_accum, "circuit blob: mix too small for the accumulators");
  B2Z_REQUIRE(c->group_size[0] >= 4 * n_accum, "circuit blob: accum group too small");

  // registers
  for (size_t i = 0; i < c->n_taps;) {
    size_t j = i;
    while (j < c->n_taps && c->tap_group[j] == c->tap_group[i] && c->tap_offset[j] == c->tap_offset[i]) j++;
    if (i > 0) {
      bool ordered = c->tap_group[i] > c->tap_group[i - 1] || (c->tap_group[i] == c->tap_group[i - 1] && c->tap_offset[i] > c->tap_offset[i - 1]);
      B2Z_REQUIRE(ordered, "circuit blob: taps not sorted by (group, offset, back)");
    }
    c->reg_group.push_back(c->tap_group[i]);
    c->reg_offset.push_back(c->tap_offset[i]);
    c->reg_tap_pos.push_back((uint32_t)i);
    c->reg_size.push_back((uint32_t)(j - i));
    // DEEP interpolates a register's evaluations at z * w^-back: duplicate backs would divide by zero (fp4_inv(0) = 0 and a
    // silently wrong U polynomial); the prover's interpolation buffers hold 16 points
    B2Z_REQUIRE(j - i <= 16, "circuit blob: a register has more than 16 taps");
    for (size_t k = i + 1; k < j; k++) B2Z_REQUIRE(c->tap_back[k] > c->tap_back[k - 1], "circuit blob: a register's backs are not strictly increasing");
    i = j;
  }
  size_t nr = c->reg_group.size();
  // every column of every group must be a register (all columns are opened by DEEP)
  {
    size_t r = 0;
    for (int g = 0; g < 3; g++)
      for (size_t col = 0; col < c->group_size[g]; col++, r++)
        B2Z_REQUIRE(r < nr && c->reg_group[r] == (uint32_t)g && c->reg_offset[r] == col, "circuit blob: a column has no tap");
    B2Z_REQUIRE(r == nr, "circuit blob: register count mismatch");
  }
  // combos in lexicographic order of their back-lists
  std::vector<uint32_t> uniq;
  for (size_t r = 0; r < nr; r++) {
    const uint32_t* b = &c->tap_back[c->reg_tap_pos[r]];
    bool found = false;
    for (uint32_t u : uniq)
      if (cmp_backs(b, c->reg_size[r], &c->tap_back[c->reg_tap_pos[u]], c->reg_size[u]) == 0) found = true;
    if (!found) uniq.push_back((uint32_t)r);
  }
  std::sort(uniq.begin(), uniq.end(), [&](uint32_t x, uint32_t y) {
    return cmp_backs(&c->tap_back[c->reg_tap_pos[x]], c->reg_size[x], &c->tap_back[c->reg_tap_pos[y]], c->reg_size[y]) < 0;
  });
  c->n_combos = uniq.size();
  size_t tot = 0;
  for (uint32_t u : uniq) {
    c->combo_begin.push_back((uint32_t)tot);
    for (uint32_t k = 0; k < c->reg_size[u]; k++) c->combo_backs.push_back(c->tap_back[c->reg_tap_pos[u] + k]);
    tot += c->reg_size[u];
  }
  c->combo_begin.push_back((uint32_t)tot);
  c->tot_combo_backs = tot;
  c->reg_combo.resize(nr);
  for (size_t r = 0; r < nr; r++)
    for (size_t k = 0; k < c->n_combos; k++)
      if (cmp_backs(&c->tap_back[c->reg_tap_pos[r]], c->reg_size[r], &c->combo_backs[c->combo_begin[k]],
                    c->combo_begin[k + 1] - c->combo_begin[k]) == 0)
        c->reg_combo[r] = (uint32_t)k;

  // static typing
  size_t nf = 0, nm = 0;
  for (const auto& s : c->steps) {
    switch (s.op) {
      case OP_CONST: case OP_GET: case OP_GET_GLOBAL: c->fp_is_ext.push_back(0); nf++; break;
      case OP_CONST_EXT: c->fp_is_ext.push_back(1); nf++; break;
      case OP_ADD: case OP_SUB: case OP_MUL:
        B2Z_REQUIRE(s.a < nf && s.b < nf, "circuit blob: forward reference");
        c->fp_is_ext.push_back(c->fp_is_ext[s.a] | c->fp_is_ext[s.b]); nf++; break;
      case OP_TRUE: c->mix_exp.push_back(0); nm++; break;
      case OP_AND_EQZ:
        B2Z_REQUIRE(s.a < nm && s.b < nf, "circuit blob: forward reference");
        c->mix_exp.push_back(c->mix_exp[s.a] + 1); nm++; break;
      case OP_AND_COND:
        B2Z_REQUIRE(s.a < nm && s.b < nf && s.c < nm, "circuit blob: forward reference");
        c->mix_exp.push_back(c->mix_exp[s.a] + c->mix_exp[s.c]); nm++; break;
      default: return set_error("circuit blob: unknown op");
    }
    if (s.op == OP_GET) B2Z_REQUIRE(s.a < c->n_taps, "circuit blob: tap index out of range");
    if (s.op == OP_GET_GLOBAL)  // operand a selects globals (0) / mix (1); the verifier and the generated kernel index both tables with b
      B2Z_REQUIRE(s.a < 2 && s.b < (s.a == 0 ? c->globals_size : c->mix_size), "circuit blob: GetGlobal operand out of range");
  }
  B2Z_REQUIRE(c->ret < nm, "circuit blob: ret out of range");
  {  // witness-generation program: base-field ops only, no forward references, operands in range
    size_t nw = 0;
    for (const auto& s : c->wit_steps) {
      switch (s.op) {
        case OP_CONST: break;
        case OP_GET: B2Z_REQUIRE(s.a < c->n_taps && c->tap_group[s.a] != B2Z_GROUP_ACCUM, "circuit blob: witgen tap out of range"); break;
        case OP_GET_GLOBAL: B2Z_REQUIRE(s.a == 0 && s.b < c->globals_size, "circuit blob: witgen GetGlobal operand out of range"); break;
        case OP_ADD: case OP_SUB: case OP_MUL: B2Z_REQUIRE(s.a < nw && s.b < nw, "circuit blob: witgen forward reference"); break;
        default: return set_error("circuit blob: op not allowed in the witgen program");
      }
      nw++;
    }
    for (size_t k = 0; k < c->wit_store_col.size(); k++)
      B2Z_REQUIRE(c->wit_store_col[k] < c->group_size[2] && c->wit_store_var[k] < nw, "circuit blob: witgen store out of range");
    for (uint32_t col : c->accum_col) B2Z_REQUIRE(col < c->group_size[2], "circuit blob: accumulator column out of range");
    for (size_t k = 0; k < c->pin_col.size(); k++)
      B2Z_REQUIRE(c->pin_col[k] < c->group_size[2] && c->pin_global[k] < c->globals_size, "circuit blob: pin out of range");
  }
  c->max_exp = 0;
  for (uint32_t e : c->mix_exp) c->max_exp = std::max<size_t>(c->max_exp, e);
  return nullptr;
}

// Seal length (words) for a given po2 -- the layout of prove/write_iop.rs + merkle.rs + fri.rs.
static void merkle_params(size_t rows, size_t& layers, size_t& top_size) {
  layers = ilog2(rows);
  top_size = (size_t)1 << pp_merkle_top_layer(layers, B2Z_QUERIES);  // [RECALL med] protocol_params.h
}
size_t circuit_seal_words(const b2z_circuit* c, unsigned po2) {
  size_t n = (size_t)1 << po2, domain = 4 * n;
  size_t words = c->globals_size + 1;
  size_t layers, top;
  merkle_params(domain, layers, top);
  size_t path = 8 * (layers - ilog2(top));
  words += 4 * 8 * top;                         // code, data, accum, check top layers
  words += 4 * (c->n_taps + B2Z_CHECK_SIZE);    // coeff_u
  size_t per_query = c->group_size[0] + c->group_size[1] + c->group_size[2] + B2Z_CHECK_SIZE + 4 * path;
  size_t size = n, dom = domain;
  while (size > B2Z_FRI_MIN_DEGREE) {
    size_t rows = dom / B2Z_FRI_FOLD, l2, t2;
    merkle_params(rows, l2, t2);
    words += 8 * t2;
    per_query += 64 + 8 * (l2 - ilog2(t2));
    size /= B2Z_FRI_FOLD;
    dom /= B2Z_FRI_FOLD;
  }
  words += 4 * size;  // final coefficients
  words += B2Z_QUERIES * per_query;
  return words;
}

// ---------------- NVRTC (dlopen'd) + cubin cache ----------------
typedef struct _nvrtcProgram* nvrtcProgram;
struct Nvrtc {
  void* h = nullptr;
  int (*CreateProgram)(nvrtcProgram*, const char*, const char*, int, const char* const*, const char* const*) = nullptr;
  int (*CompileProgram)(nvrtcProgram, int, const char* const*) = nullptr;
  int (*GetCUBINSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetCUBIN)(nvrtcProgram, char*) = nullptr;
  int (*GetProgramLogSize)(nvrtcProgram, size_t*) = nullptr;
  int (*GetProgramLog)(nvrtcProgram, char*) = nullptr;
  int (*DestroyProgram)(nvrtcProgram*) = nullptr;
  int (*Version)(int*, int*) = nullptr;
  bool load() {
    if (h) return true;
    const char* names[] = {"libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so.12", "/usr/local/cuda/lib64/libnvrtc.so", "libnvrtc.so"};
    for (const char* nme : names) {
      h = dlopen(nme, RTLD_NOW | RTLD_LOCAL);
      if (h) break;
    }
    if (!h) return false;
#define LOADSYM(field, name) *(void**)(&field) = dlsym(h, name); if (!field) return false;
    LOADSYM(CreateProgram, "nvrtcCreateProgram")
    LOADSYM(CompileProgram, "nvrtcCompileProgram")
    LOADSYM(GetCUBINSize, "nvrtcGetCUBINSize")
    LOADSYM(GetCUBIN, "nvrtcGetCUBIN")
    LOADSYM(GetProgramLogSize, "nvrtcGetProgramLogSize")
    LOADSYM(GetProgramLog, "nvrtcGetProgramLog")
    LOADSYM(DestroyProgram, "nvrtcDestroyProgram")
    LOADSYM(Version, "nvrtcVersion")
#undef LOADSYM
    return true;
  }
};
static Nvrtc g_nvrtc;


// ---------------- code generation ----------------
static const char* PRELUDE = R"CUDA(
// generated by libb2z circuit_codegen -- BabyBear field prelude (same reduction sequence as csrc/field.cuh)
typedef unsigned int u32;
typedef unsigned long long u64;
#define P 2013265921u
#define PINV 0x88000001u
#define BETA 0x37ffffe9u
__device__ __forceinline__ u32 fadd(u32 a, u32 b) { u32 r = a + b; return min(r, r - P); }
__device__ __forceinline__ u32 fsub(u32 a, u32 b) { u32 r = a - b; return min(r, r + P); }
__device__ __forceinline__ u32 fmul(u32 a, u32 b) {
  u64 t = (u64)a * b; u32 m = (u32)t * PINV; u32 u = __umulhi(m, P); u32 r = (u32)(t >> 32) - u; return min(r, r + P);
}
// sums of up to four products reduced ONCE: every product of two canonical words is < P^2 and 4 P^2 < 2^64, so the terms are
// accumulated in 64 bits (one IMAD.WIDE each) and the sum x is brought to x * 2^-32 mod P by one Montgomery step
__device__ __forceinline__ u32 red64(u64 x) {
  u32 hi = (u32)(x >> 32), lo = (u32)x;
  hi = min(hi, hi - P); hi = min(hi, hi - P);  // hi < 2^32 < 3 P -> hi < P
  u32 m = lo * PINV; u32 u = __umulhi(m, P); u32 r = hi - u; return min(r, r + P);
}
__device__ __forceinline__ u32 dot2(u32 a0, u32 b0, u32 a1, u32 b1) { return red64((u64)a0 * b0 + (u64)a1 * b1); }
__device__ __forceinline__ u32 dot3(u32 a0, u32 b0, u32 a1, u32 b1, u32 a2, u32 b2) { return red64((u64)a0 * b0 + (u64)a1 * b1 + (u64)a2 * b2); }
__device__ __forceinline__ u32 dot4(u32 a0, u32 b0, u32 a1, u32 b1, u32 a2, u32 b2, u32 a3, u32 b3) {
  return red64((u64)a0 * b0 + (u64)a1 * b1 + (u64)a2 * b2 + (u64)a3 * b3);
}
struct E { u32 a, b, c, d; };
__device__ __forceinline__ E emake(u32 a, u32 b, u32 c, u32 d) { E r; r.a = a; r.b = b; r.c = c; r.d = d; return r; }
__device__ __forceinline__ E eadd(E x, E y) { return emake(fadd(x.a, y.a), fadd(x.b, y.b), fadd(x.c, y.c), fadd(x.d, y.d)); }
__device__ __forceinline__ E esub(E x, E y) { return emake(fsub(x.a, y.a), fsub(x.b, y.b), fsub(x.c, y.c), fsub(x.d, y.d)); }
__device__ __forceinline__ E eaddf(E x, u32 y) { return emake(fadd(x.a, y), x.b, x.c, x.d); }
__device__ __forceinline__ E esubf(E x, u32 y) { return emake(fsub(x.a, y), x.b, x.c, x.d); }
__device__ __forceinline__ E fsube(u32 x, E y) { return emake(fsub(x, y.a), fsub(0u, y.b), fsub(0u, y.c), fsub(0u, y.d)); }
__device__ __forceinline__ E emulf(E x, u32 y) { return emake(fmul(x.a, y), fmul(x.b, y), fmul(x.c, y), fmul(x.d, y)); }
__device__ __forceinline__ E emul(E x, E y) {
  E r;
  r.a = fadd(fmul(x.a, y.a), fmul(BETA, fadd(fadd(fmul(x.b, y.d), fmul(x.c, y.c)), fmul(x.d, y.b))));
  r.b = fadd(fadd(fmul(x.a, y.b), fmul(x.b, y.a)), fmul(BETA, fadd(fmul(x.c, y.d), fmul(x.d, y.c))));
  r.c = fadd(fadd(fadd(fmul(x.a, y.c), fmul(x.b, y.b)), fmul(x.c, y.a)), fmul(BETA, fmul(x.d, y.d)));
  r.d = fadd(fadd(fmul(x.a, y.d), fmul(x.b, y.c)), fadd(fmul(x.c, y.b), fmul(x.d, y.a)));
  return r;
}
// loads are volatile asm so that the compiler keeps each tap load next to its use instead of hoisting hundreds of
// loads to the top of the straight-line program (which costs 255 registers and spills)
__device__ __forceinline__ u32 ldv(const u32* p) { u32 v; asm volatile("ld.global.nc.u32 %0, [%1];" : "=r"(v) : "l"(p)); return v; }
// column load: base pointer of the column (kernel-parameter table, uniform) + 32-bit byte offset of the row
__device__ __forceinline__ u32 ldc(const u32* col, u32 off) {
  u32 v;
  asm volatile("{ .reg .u64 a; cvt.u64.u32 a, %2; add.u64 a, a, %1; ld.global.nc.u32 %0, [a]; }" : "=r"(v) : "l"(col), "r"(off));
  return v;
}
__device__ __forceinline__ E eld(const u32* p) {
  E r; asm volatile("ld.global.nc.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(r.a), "=r"(r.b), "=r"(r.c), "=r"(r.d) : "l"(p)); return r;
}
)CUDA";

// Straight-line emission with component-wise constant propagation.
//   * every value is 1 (base field) or 4 (extension) scalar operands; an operand is a compile-time constant (Montgomery
//     form) or the name of a `const u32` temporary / constant-bank parameter;
//   * products and sums fold constants: x*0 = 0, x*1 = x, x+0 = x.  Programs in this vocabulary build extension values
//     as sum_k e_k * x_k (ConstExt basis vectors) -- folding turns those 16-product ext multiplications into moves;
//   * leaf steps (Get) are re-materialised as a fresh load at every use, so a tap value lives for one instruction.
struct Opd {
  bool is_const = true;
  uint32_t k = 0;     // Montgomery constant when is_const
  int tid = -1;       // id of the temporary when the operand is one (statement collection / partitioning)
  std::string name;   // expression otherwise
  std::string str() const { return is_const ? std::to_string(k) + "u" : name; }
};
struct Val {
  bool ext = false;
  Opd c[4];
  bool leaf_load = false;  // re-emit the load at each use
  std::string load_expr;
  std::string load_col, load_off;  // column pointer / row offset expressions of a tap load (for staging)
};
struct Emitter {
  const b2z_circuit* c;
  bool trace_domain;  // witgen: taps address the trace (stride 1) instead of the 4x LDE (stride 4)
  std::ostringstream* o;
  std::vector<Val> vals;
  unsigned tmp = 0;
  static Opd konst(uint32_t k) { Opd r; r.is_const = true; r.k = k; return r; }
  static Opd named(const std::string& n) { Opd r; r.is_const = false; r.name = n; return r; }
  // collect mode: statements are recorded (with the temporaries they read) instead of printed, so that the program can
  // be cut into __noinline__ parts afterwards
  struct Stmt {
    unsigned id;
    std::string expr;
    std::vector<unsigned> uses;
  };
  std::vector<Stmt>* stmts = nullptr;
  std::string cp_name = "cp", kc_name = "kc", glob_name = "globals";
  // tiled mode (eval_check over the tiled LDE copy): a tap is `ldt(pt + constant)`; flat column index = tile_flat0[group] + column
  bool tiled = false;
  size_t tile_flat0[3] = {0, 0, 0};
  unsigned tile_pitch_bytes = 0;
  Opd emit(const std::string& expr, const Opd* a = nullptr, const Opd* b = nullptr) {
    unsigned id = tmp++;
    std::string n = "t" + std::to_string(id);
    if (stmts) {
      Stmt st{id, expr, {}};
      if (a && a->tid >= 0) st.uses.push_back((unsigned)a->tid);
      if (b && b->tid >= 0) st.uses.push_back((unsigned)b->tid);
      stmts->push_back(st);
    } else {
      *o << "  const u32 " << n << " = " << expr << ";\n";
    }
    Opd r = named(n);
    r.tid = (int)id;
    return r;
  }
  Opd emit_n(const std::string& expr, const std::vector<Opd>& ops) {
    unsigned id = tmp++;
    std::string n = "t" + std::to_string(id);
    if (stmts) {
      Stmt st{id, expr, {}};
      for (const auto& x : ops)
        if (x.tid >= 0) st.uses.push_back((unsigned)x.tid);
      stmts->push_back(st);
    } else {
      *o << "  const u32 " << n << " = " << expr << ";\n";
    }
    Opd r = named(n);
    r.tid = (int)id;
    return r;
  }
  Opd s_mul(const Opd& a, const Opd& b) {
    if (a.is_const && b.is_const) return konst(fp_mul(a.k, b.k));
    if (a.is_const) {
      if (a.k == 0) return konst(0);
      if (a.k == MONT_ONE) return b;
    }
    if (b.is_const) {
      if (b.k == 0) return konst(0);
      if (b.k == MONT_ONE) return a;
    }
    return emit("fmul(" + a.str() + ", " + b.str() + ")", &a, &b);
  }
  Opd s_add(const Opd& a, const Opd& b) {
    if (a.is_const && b.is_const) return konst(fp_add(a.k, b.k));
    if (a.is_const && a.k == 0) return b;
    if (b.is_const && b.k == 0) return a;
    return emit("fadd(" + a.str() + ", " + b.str() + ")", &a, &b);
  }
  Opd s_sub(const Opd& a, const Opd& b) {
    if (a.is_const && b.is_const) return konst(fp_sub(a.k, b.k));
    if (b.is_const && b.k == 0) return a;
    return emit("fsub(" + a.str() + ", " + b.str() + ")", &a, &b);
  }
  // operand view of a value; leaves are loaded here (fresh temporary per use)
  Val use(size_t v) {
    Val r = vals[v];
    if (r.leaf_load) {
      r.c[0] = emit(r.load_expr);
      r.leaf_load = false;
    }
    return r;
  }
  static Val base(const Opd& x) {
    Val r;
    r.ext = false;
    r.c[0] = x;
    return r;
  }
  Val v_add(const Val& x, const Val& y, bool sub) {
    Val r;
    r.ext = x.ext | y.ext;
    int n = r.ext ? 4 : 1;
    for (int i = 0; i < n; i++) {
      Opd a = (i == 0 || x.ext) ? x.c[i] : konst(0);
      Opd b = (i == 0 || y.ext) ? y.c[i] : konst(0);
      r.c[i] = sub ? s_sub(a, b) : s_add(a, b);
    }
    return r;
  }
  Val v_mul(const Val& x, const Val& y) {
    Val r;
    r.ext = x.ext | y.ext;
    if (!x.ext && !y.ext) {
      r.c[0] = s_mul(x.c[0], y.c[0]);
    } else if (x.ext != y.ext) {
      const Val& e = x.ext ? x : y;
      const Opd& b = x.ext ? y.c[0] : x.c[0];
      for (int i = 0; i < 4; i++) r.c[i] = s_mul(e.c[i], b);
    } else {
      const Opd beta = konst(MONT_BETA);
      const Opd *a = x.c, *b = y.c;
      r.c[0] = s_add(s_mul(a[0], b[0]), s_mul(beta, s_add(s_add(s_mul(a[1], b[3]), s_mul(a[2], b[2])), s_mul(a[3], b[1]))));
      r.c[1] = s_add(s_add(s_mul(a[0], b[1]), s_mul(a[1], b[0])), s_mul(beta, s_add(s_mul(a[2], b[3]), s_mul(a[3], b[2]))));
      r.c[2] = s_add(s_add(s_add(s_mul(a[0], b[2]), s_mul(a[1], b[1])), s_mul(a[2], b[0])), s_mul(beta, s_mul(a[3], b[3])));
      r.c[3] = s_add(s_add(s_mul(a[0], b[3]), s_mul(a[1], b[2])), s_add(s_mul(a[2], b[1]), s_mul(a[3], b[0])));
    }
    return r;
  }
  void step(const b2z_circuit::Step& s) {
    Val r;
    switch (s.op) {
      case OP_CONST:
        r = base(konst(fp_from_u32(s.a)));
        break;
      case OP_CONST_EXT:
        r.ext = true;
        r.c[0] = konst(fp_from_u32(s.a));
        r.c[1] = konst(fp_from_u32(s.b));
        r.c[2] = konst(fp_from_u32(s.c));
        r.c[3] = konst(fp_from_u32(s.d));
        break;
      case OP_GET: {
        uint32_t g = c->tap_group[s.a], col = c->tap_offset[s.a], back = c->tap_back[s.a];
        r.ext = false;
        r.leaf_load = true;
        r.load_col = cp_name + ".g" + std::to_string(g) + "[" + std::to_string(col) + "]";
        r.load_off = "ib" + std::to_string(back);
        r.load_expr = "ldc(" + r.load_col + ", " + r.load_off + ")";
        if (tiled) {
          const long long off = (long long)(tile_flat0[g] + col) * tile_pitch_bytes - 16ll * back;
          r.load_expr = "ldt(pt + (" + std::to_string(off) + "))";
        }
      } break;
      case OP_GET_GLOBAL:
        if (trace_domain) {
          r.ext = false;
          r.leaf_load = true;
          r.load_expr = "ldv(" + glob_name + " + " + std::to_string(s.b) + ")";
        } else {
          r = base(named(kc_name + (s.a == 0 ? ".g[" : ".m[") + std::to_string(s.b) + "]"));
        }
        break;
      case OP_ADD: r = v_add(use(s.a), use(s.b), false); break;
      case OP_SUB: r = v_add(use(s.a), use(s.b), true); break;
      case OP_MUL: r = v_mul(use(s.a), use(s.b)); break;
      default: break;
    }
    vals.push_back(r);
  }
};

static Val pw_val(uint32_t e, const std::string& kc_name = "kc") {
  Val r;
  r.ext = true;
  if (e == 0) {
    r.c[0] = Emitter::konst(MONT_ONE);
    for (int i = 1; i < 4; i++) r.c[i] = Emitter::konst(0);
  } else {
    for (int i = 0; i < 4; i++) r.c[i] = Emitter::named(kc_name + ".pw[" + std::to_string(4 * e + i) + "]");
  }
  return r;
}


// ---------------- staged eval_check: the tile's columns live in shared memory ----------------
// The per-use tap loads of the straight-line program are its largest cost at rv32im scale: 18.6 k loads per point, 3 address
// instructions each, 74 KB per point through L1 / L2 (ncu: 16.5 GB of DRAM reads for a 1 GB working set at po2 18).  Here a CTA
// owns a tile of T = 128 consecutive coset points and first stages EVERY column of the three LDE matrices for that tile (plus a
// halo of 4 * max_back rows for the columns that are tapped at earlier rows) into shared memory -- 144 KB for W = 256; every tap
// of every point is then ONE `ld.shared` with a compile-time offset.  To fill the SM with one such tile resident, the constraint
// program is split: the longest AndEqz / AndCond chain (a SUM of terms: tot(chain) = sum_k alpha^e_k * cond_k * value_k) is cut into
// G contiguous ranges, thread group g (T threads) evaluates range g for all T points, the partial sums meet in shared memory and
// group 0 runs the few steps after the chain.  Field arithmetic is exact, so the regrouped sum is bit-identical.
struct DemandEmitter : Emitter {
  std::vector<size_t> fp_def, mix_def;     // variable -> defining step
  std::vector<char> have;
  std::vector<uint32_t> tap_byte_off;      // per tap: byte offset of (this thread's row, this column, this back) from the thread's base
  std::vector<size_t>* cut_points = nullptr;
  struct MixAcc {
    Val tot;
    std::vector<std::pair<uint32_t, Opd>> pend;
  };
  std::map<size_t, MixAcc> mix_memo;
  Val need(size_t v) {
    if (!have[v]) {
      const b2z_circuit::Step& s = c->steps[fp_def[v]];
      Val r;
      switch (s.op) {
        case OP_CONST: r = base(konst(fp_from_u32(s.a))); break;
        case OP_CONST_EXT:
          r.ext = true;
          r.c[0] = konst(fp_from_u32(s.a));
          r.c[1] = konst(fp_from_u32(s.b));
          r.c[2] = konst(fp_from_u32(s.c));
          r.c[3] = konst(fp_from_u32(s.d));
          break;
        case OP_GET:
          r.leaf_load = true;
          r.load_expr = "lds(tp + " + std::to_string(tap_byte_off[s.a]) + "u)";
          break;
        case OP_GET_GLOBAL: r = base(named(kc_name + (s.a == 0 ? ".g[" : ".m[") + std::to_string(s.b) + "]")); break;
        case OP_ADD: { Val x = need(s.a), y = need(s.b); r = v_add(x, y, false); } break;
        case OP_SUB: { Val x = need(s.a), y = need(s.b); r = v_add(x, y, true); } break;
        case OP_MUL: { Val x = need(s.a), y = need(s.b); r = v_mul(x, y); } break;
        default: break;
      }
      vals[v] = r;
      have[v] = 1;
    }
    return use(v);
  }
  void flush(MixAcc& m) {
    if (m.pend.empty()) return;
    for (int k = 0; k < 4; k++) {
      std::string expr;
      std::vector<Opd> ops;
      size_t nterm = 0;
      for (auto& pr : m.pend) {
        if (pr.first == 0) continue;
        expr += (nterm ? ", " : "") + kc_name + ".pw[" + std::to_string(4 * pr.first + k) + "], " + pr.second.str();
        ops.push_back(pr.second);
        nterm++;
      }
      if (nterm == 1) m.tot.c[k] = s_add(m.tot.c[k], emit_n("fmul(" + expr + ")", ops));
      else if (nterm >= 2) m.tot.c[k] = s_add(m.tot.c[k], emit_n("dot" + std::to_string(nterm) + "(" + expr + ")", ops));
    }
    for (auto& pr : m.pend)
      if (pr.first == 0) m.tot.c[0] = s_add(m.tot.c[0], pr.second);
    m.pend.clear();
  }
  // total of mix variable m (iterative along its chain, recursive only into AndCond's inner chains)
  MixAcc need_mix(size_t m) {
    std::vector<size_t> path;
    size_t cur = m;
    while (!mix_memo.count(cur)) {
      const b2z_circuit::Step& s = c->steps[mix_def[cur]];
      if (s.op == OP_TRUE) {
        MixAcc z;
        z.tot.ext = true;
        mix_memo[cur] = z;
        break;
      }
      path.push_back(cur);
      cur = s.a;
    }
    for (auto it = path.rbegin(); it != path.rend(); ++it) {
      const b2z_circuit::Step& s = c->steps[mix_def[*it]];
      MixAcc n = mix_memo[s.a];
      if (s.op == OP_AND_EQZ) {
        Val v = need(s.b);
        if (!v.ext && !v.c[0].is_const) {
          n.pend.push_back({c->mix_exp[s.a], v.c[0]});
          if (n.pend.size() == 4) flush(n);
        } else {
          flush(n);
          n.tot = v_add(n.tot, v_mul(pw_val(c->mix_exp[s.a], kc_name), v), false);
        }
      } else {  // OP_AND_COND
        MixAcc in = need_mix(s.c);
        flush(n);
        flush(in);
        Val cond = need(s.b);
        Val inner = v_mul(in.tot, cond);
        n.tot = v_add(n.tot, v_mul(inner, pw_val(c->mix_exp[s.a], kc_name)), false);
      }
      mix_memo[*it] = n;
      if (cut_points) cut_points->push_back(stmts->size());
    }
    return mix_memo[m];
  }
};

// cut a statement list into __noinline__ parts (same rule as the unstaged emitter); returns the calling sequence
static std::string emit_staged_parts(std::ostringstream& o, const std::string& prefix, const std::vector<Emitter::Stmt>& stmts,
                                     const std::vector<size_t>& cut_points, size_t parts_m, const std::vector<Opd>& live_out,
                                     const std::string& extra_params, const std::string& extra_args) {
  std::vector<size_t> bounds{0};
  for (size_t cp_i : cut_points)
    if (cp_i >= bounds.back() + parts_m && cp_i <= stmts.size()) bounds.push_back(cp_i);
  if (bounds.back() != stmts.size()) bounds.push_back(stmts.size());
  std::map<unsigned, size_t> def_part, last_use_part;
  for (size_t p = 0; p + 1 < bounds.size(); p++)
    for (size_t k = bounds[p]; k < bounds[p + 1]; k++) {
      def_part[stmts[k].id] = p;
      for (unsigned u : stmts[k].uses) last_use_part[u] = p;
    }
  const size_t FINAL = bounds.size();
  for (const auto& x : live_out)
    if (x.tid >= 0) last_use_part[(unsigned)x.tid] = FINAL;
  std::ostringstream body;
  for (size_t p = 0; p + 1 < bounds.size(); p++) {
    std::vector<unsigned> ins, outs;
    std::map<unsigned, bool> seen;
    for (size_t k = bounds[p]; k < bounds[p + 1]; k++) {
      for (unsigned u : stmts[k].uses)
        if (def_part[u] < p && !seen[u]) {
          seen[u] = true;
          ins.push_back(u);
        }
      if (last_use_part.count(stmts[k].id) && last_use_part[stmts[k].id] > p) outs.push_back(stmts[k].id);
    }
    o << "struct " << prefix << "R" << p << " { u32 v[" << (outs.empty() ? 1 : outs.size()) << "]; };\n";
    o << "static __device__ __noinline__ " << prefix << "R" << p << " " << prefix << "p" << p << "(" << extra_params;
    for (unsigned u : ins) o << ", u32 t" << u;
    o << ") {\n";
    for (size_t k = bounds[p]; k < bounds[p + 1]; k++) o << "  const u32 t" << stmts[k].id << " = " << stmts[k].expr << ";\n";
    o << "  " << prefix << "R" << p << " r;\n";
    if (outs.empty()) o << "  r.v[0] = 0u;\n";
    for (size_t j = 0; j < outs.size(); j++) o << "  r.v[" << j << "] = t" << outs[j] << ";\n";
    o << "  return r;\n}\n";
    body << "    const " << prefix << "R" << p << " " << prefix << "r" << p << " = " << prefix << "p" << p << "(" << extra_args;
    for (unsigned u : ins) body << ", t" << u;
    body << ");\n";
    for (size_t j = 0; j < outs.size(); j++) body << "    const u32 t" << outs[j] << " = " << prefix << "r" << p << ".v[" << j << "];\n";
  }
  return body.str();
}

static bool codegen_eval_staged(b2z_circuit* c, std::ostringstream& o, int parts_m) {
  constexpr unsigned T = 128;
  int G = 8;
  if (const char* e = getenv("B2Z_EVALCHECK_GROUPS")) G = atoi(e);
  if (G < 1 || G > 8) G = 8;
  const size_t ws[3] = {c->group_size[0] ? c->group_size[0] : 1, c->group_size[1] ? c->group_size[1] : 1, c->group_size[2] ? c->group_size[2] : 1};
  const size_t flat0[3] = {0, ws[0], ws[0] + ws[1]};
  const size_t ncol = ws[0] + ws[1] + ws[2];
  // staging layout: column (g, col) keeps rows [i0 - halo, i0 + T), halo = 4 * deepest back of the column
  std::vector<uint32_t> halo(ncol, 0), off(ncol, 0);
  for (size_t t = 0; t < c->n_taps; t++) {
    size_t f = flat0[c->tap_group[t]] + c->tap_offset[t];
    halo[f] = std::max(halo[f], 4 * c->tap_back[t]);
  }
  size_t words = 0;
  for (size_t f = 0; f < ncol; f++) {
    off[f] = (uint32_t)words;
    words += T + halo[f];
  }
  // variable -> defining step; the main chain = the longest AndEqz/AndCond chain
  DemandEmitter proto{};
  std::vector<size_t> fp_def, mix_def;
  for (size_t i = 0; i < c->steps.size(); i++) {
    uint32_t op = c->steps[i].op;
    if (op == OP_TRUE || op == OP_AND_EQZ || op == OP_AND_COND) mix_def.push_back(i);
    else fp_def.push_back(i);
  }
  const size_t nm = mix_def.size();
  // The chain to split: start at `ret`; while ONE AndCond link of the current chain owns most of the constraints below it
  // (weight = number of AndEqz / AndCond steps, nested chains included), descend into that link's inner chain.
  std::vector<size_t> weight(nm, 0);
  for (size_t m = 0; m < nm; m++) {
    const auto& st = c->steps[mix_def[m]];
    if (st.op != OP_TRUE) weight[m] = weight[st.a] + 1 + (st.op == OP_AND_COND ? weight[st.c] : 0);
  }
  size_t X = c->ret;
  for (;;) {
    size_t heavy = nm, heavy_w = 0;
    for (size_t m = X; c->steps[mix_def[m]].op != OP_TRUE; m = c->steps[mix_def[m]].a) {
      const auto& st = c->steps[mix_def[m]];
      if (st.op == OP_AND_COND && weight[st.c] + 1 > heavy_w) {
        heavy_w = weight[st.c] + 1;
        heavy = m;
      }
    }
    if (heavy == nm || 10 * heavy_w < 6 * weight[X]) break;
    X = c->steps[mix_def[heavy]].c;
  }
  std::vector<size_t> chain;  // variables of the main chain, in order (without its True)
  for (size_t m = X; c->steps[mix_def[m]].op != OP_TRUE; m = c->steps[mix_def[m]].a) chain.push_back(m);
  std::reverse(chain.begin(), chain.end());
  const size_t chain_true = chain.empty() ? X : c->steps[mix_def[chain.front()]].a;
  std::vector<char> on_chain(nm, 0);
  for (size_t m : chain) on_chain[m] = 1;
  // only X may be referenced from outside the chain's own links (otherwise the split would need intermediate totals)
  for (size_t m = 0; m < nm; m++) {
    const auto& s = c->steps[mix_def[m]];
    if (s.op == OP_TRUE) continue;
    const bool next_link = on_chain[m];
    if (!next_link && s.a != X && (on_chain[s.a] || (s.a == chain_true && !chain.empty() && false))) return false;
    if (s.op == OP_AND_COND && on_chain[s.c] && s.c != X) return false;
    if (next_link && s.op == OP_AND_COND && on_chain[s.c]) return false;
  }
  if ((int)chain.size() < 2 * G) G = 1;
  const size_t smem_bytes = 4 * (words + 4 * (size_t)G * T);
  if (smem_bytes > 220 * 1024) return false;  // wider circuits: fall back to per-use global loads
  // tap -> byte offset from the thread's base (shared address of word `t` of the tile area)
  std::vector<uint32_t> tap_byte_off(c->n_taps);
  for (size_t t = 0; t < c->n_taps; t++) {
    size_t f = flat0[c->tap_group[t]] + c->tap_offset[t];
    tap_byte_off[t] = 4 * (off[f] + halo[f] - 4 * c->tap_back[t]);
  }
  auto make = [&](std::vector<Emitter::Stmt>& st, std::vector<size_t>& cuts) {
    DemandEmitter em{};
    em.c = c;
    em.trace_domain = false;
    em.o = &o;
    em.stmts = &st;
    em.kc_name = "ckc";
    em.fp_def = fp_def;
    em.mix_def = mix_def;
    em.have.assign(fp_def.size(), 0);
    em.vals.assign(fp_def.size(), Val());
    em.tap_byte_off = tap_byte_off;
    em.cut_points = &cuts;
    return em;
  };
  // group ranges: equal numbers of generated statements (measured by a dry run over the whole chain)
  std::vector<size_t> rng(G + 1, chain.size());
  rng[0] = 0;
  if (G > 1) {
    std::vector<size_t> cum(chain.size(), 0);
    {
      std::vector<Emitter::Stmt> st;
      std::vector<size_t> cuts;
      DemandEmitter em = make(st, cuts);
      DemandEmitter::MixAcc zero;
      zero.tot.ext = true;
      em.mix_memo[chain_true] = zero;
      for (size_t k = 0; k < chain.size(); k++) {
        em.need_mix(chain[k]);
        cum[k] = st.size();
      }
    }
    for (int g = 1; g < G; g++) {
      const size_t target = cum.back() * g / G;
      size_t k = rng[g - 1];
      while (k < chain.size() && cum[k] < target) k++;
      rng[g] = std::max(k, rng[g - 1]);
    }
  }
  o << "__constant__ ColPtrs ccp;\n__constant__ Consts ckc;\n";
  o << "__device__ __forceinline__ u32 lds(u32 a) { u32 v; asm volatile(\"ld.shared.u32 %0, [%1];\" : \"=r\"(v) : \"r\"(a)); return v; }\n";
  o << "__constant__ u32 stg_off[" << ncol << "] = {";
  for (size_t f = 0; f < ncol; f++) o << (f ? ", " : "") << off[f];
  o << "};\n__constant__ u32 stg_halo[" << ncol << "] = {";
  for (size_t f = 0; f < ncol; f++) o << (f ? ", " : "") << halo[f];
  o << "};\n";
  o << "// staging: warp w copies columns w, w + nwarps, ...: rows [i0 - halo, i0 + " << T << ") of the column, indices wrapped to the domain\n";
  o << "__device__ __forceinline__ void stage_tile(u32* st, u64 i0, u64 mask) {\n"
       "  const unsigned warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;\n"
       "  const u32* const* cols = reinterpret_cast<const u32* const*>(&ccp);\n"
       "  for (unsigned col = warp; col < "
    << ncol << "u; col += nw) {\n"
       "    const u32* src = cols[col];\n    const unsigned h = stg_halo[col], n = "
    << T << "u + h;\n    u32* dst = st + stg_off[col];\n"
       "    for (unsigned r = lane; r < n; r += 32) dst[r] = __ldg(src + ((i0 - h + r) & mask));\n  }\n}\n";
  // groups
  std::vector<std::string> bodies(G);
  std::vector<Val> partial(G);
  for (int g = 0; g < G; g++) {
    std::vector<Emitter::Stmt> st;
    std::vector<size_t> cuts;
    DemandEmitter em = make(st, cuts);
    DemandEmitter::MixAcc zero;
    zero.tot.ext = true;
    Val tot = zero.tot;
    if (rng[g + 1] > rng[g]) {
      em.mix_memo[rng[g] == 0 ? chain_true : chain[rng[g] - 1]] = zero;
      DemandEmitter::MixAcc r = em.need_mix(chain[rng[g + 1] - 1]);
      em.flush(r);
      tot = r.tot;
    }
    partial[g] = tot;
    std::vector<Opd> live{tot.c[0], tot.c[1], tot.c[2], tot.c[3]};
    bodies[g] = emit_staged_parts(o, "g" + std::to_string(g), st, cuts, (size_t)parts_m, live, "u32 tp", "tp");
  }
  // epilogue: the steps after the main chain, fed with the summed partial totals
  std::string epi_body;
  Val result;
  {
    std::vector<Emitter::Stmt> st;
    std::vector<size_t> cuts;
    DemandEmitter em = make(st, cuts);
    em.tmp = 1u << 30;  // temporaries of the epilogue must not collide with s0..s3 / the groups' names
    DemandEmitter::MixAcc xs;
    xs.tot.ext = true;
    for (int k = 0; k < 4; k++) xs.tot.c[k] = Emitter::named("s" + std::to_string(k));
    if (!chain.empty()) em.mix_memo[X] = xs;
    DemandEmitter::MixAcc r = em.need_mix(c->ret);
    em.flush(r);
    result = r.tot;
    std::vector<Opd> live{result.c[0], result.c[1], result.c[2], result.c[3]};
    epi_body = emit_staged_parts(o, "e", st, cuts, (size_t)parts_m, live, "u32 tp, u32 s0, u32 s1, u32 s2, u32 s3", "tp, s0, s1, s2, s3");
  }
  o << "extern \"C\" __global__ void __launch_bounds__(" << G * T << ", 1) b2z_eval_check(u32* __restrict__ check, u32 iz0, u32 iz1, u32 iz2, u32 iz3, u64 mask) {\n";
  o << "  extern __shared__ u32 st[];\n  const u64 domain = mask + 1;\n";
  o << "  const unsigned t = threadIdx.x & " << (T - 1) << "u, grp = threadIdx.x >> 7;\n";
  o << "  const u64 i0 = (u64)blockIdx.x * " << T << "ull;\n  stage_tile(st, i0, mask);\n  __syncthreads();\n";
  o << "  const u32 tp = (u32)__cvta_generic_to_shared(st) + 4u * t;\n  u32 p0 = 0u, p1 = 0u, p2 = 0u, p3 = 0u;\n";
  for (int g = 0; g < G; g++) {
    o << (g ? "  else if" : "  if") << " (grp == " << g << "u) {\n" << bodies[g];
    for (int k = 0; k < 4; k++) o << "    p" << k << " = " << partial[g].c[k].str() << ";\n";
    o << "  }\n";
  }
  o << "  u32* ps = st + " << words << "u;  // partial totals [4][" << G << "][" << T << "]\n";
  for (int k = 0; k < 4; k++) o << "  ps[(" << k << "u * " << G << "u + grp) * " << T << "u + t] = p" << k << ";\n";
  o << "  __syncthreads();\n  if (grp != 0u) return;\n";
  for (int k = 0; k < 4; k++) {
    o << "  u32 s" << k << " = ps[" << k * G * T << "u + t];\n";
    for (int g = 1; g < G; g++) o << "  s" << k << " = fadd(s" << k << ", ps[" << (k * G + g) * T << "u + t]);\n";
  }
  o << "  {\n" << epi_body << "    const u64 i = i0 + t;\n    if (i > mask) return;\n";
  o << "    const u32 iz = (i & 2) ? ((i & 1) ? iz3 : iz2) : ((i & 1) ? iz1 : iz0);\n";
  for (int k = 0; k < 4; k++) {
    std::string v = (result.ext || k == 0) ? result.c[k].str() : std::string("0u");
    o << "    check[" << k << " * domain + i] = fmul(" << v << ", iz);\n";
  }
  o << "  }\n}\n\n";
  c->parts_mode = true;
  c->eval_threads = (unsigned)(G * T);
  c->eval_tiles = 1;
  c->points_per_thread = 1;
  c->eval_points_per_cta = T;
  c->eval_smem_bytes = smem_bytes;
  return true;
}

void circuit_codegen(b2z_circuit* c) {
  std::ostringstream o;
  o << PRELUDE;
  o << "// circuit \"" << std::string((const char*)c->info, 16) << "\": " << c->steps.size() << " constraint steps, " << c->n_taps
    << " taps\n";
  // ---- eval_check: one thread per point of the 4N coset ----
  // column base pointers travel as one by-value kernel parameter (constant bank => uniform registers, no per-load
  // 64-bit multiply); rows are addressed with 32-bit byte offsets ib<back> = ((i - 4*back) & mask) * 4
  o << "struct ColPtrs { const u32* g0[" << (c->group_size[0] ? c->group_size[0] : 1) << "]; const u32* g1["
    << (c->group_size[1] ? c->group_size[1] : 1) << "]; const u32* g2[" << (c->group_size[2] ? c->group_size[2] : 1) << "]; };\n";
  unsigned max_back = 0;
  for (size_t t = 0; t < c->n_taps; t++) max_back = std::max(max_back, c->tap_back[t]);
  // register cap = 65536 / (128 * min_blocks); tunable for experiments through B2Z_EVALCHECK_MINBLOCKS
  int min_blocks = 9;
  if (const char* e = getenv("B2Z_EVALCHECK_MINBLOCKS")) min_blocks = atoi(e) > 0 ? atoi(e) : 9;
  // globals, mix values and the powers of poly_mix are launch constants: they travel by value (constant bank), so
  // they are instruction operands and cost neither loads nor registers
  o << "struct Consts { u32 g[" << (c->globals_size ? c->globals_size : 1) << "]; u32 m[" << (c->mix_size ? c->mix_size : 1)
    << "]; u32 pw[" << 4 * (c->max_exp + 1) << "]; };\n";
  // B2Z_EVALCHECK_PARTS = M > 0: cut the straight-line program into __noinline__ device functions of about M statements
  // (cuts right after a constraint has been accumulated, so only the accumulator and the few shared temporaries cross a
  // cut).  The generated CUDA of upstream's rv32im circuit is split the same way; ptxas cannot hoist loads across calls,
  // which is what makes one huge basic block spill.  Tables then live in module-level __constant__ memory.
  // Measured on B200 (planning circuit, po2 20): straight line 3.79 ms (64 regs + 312 B spills), parts of 60 / 120 / 250 /
  // 500 statements 2.71 / 2.63 / 2.68 / 2.86 ms, no spills.  Default 120 (B2Z_EVALCHECK_PARTS=0 restores the single block).
  int parts_m = 120;
  if (const char* e = getenv("B2Z_EVALCHECK_PARTS")) parts_m = atoi(e);
  c->parts_mode = parts_m > 0;
  // B2Z_EVALCHECK_POINTS = K > 1 (parts mode only): every thread evaluates K coset points and calls each part K times in
  // a row, so a part's instructions are fetched once per K points.  The straight-line program is far larger than the
  // instruction cache (planning circuit: ~190 KB of SASS, ncu: 49.5 % i-cache hits, "no instruction" the top stall);
  // only the accumulator (4..8 words per point) is live across parts, so K points cost K x 8 registers.
  int points_k = 1;
  if (const char* e = getenv("B2Z_EVALCHECK_POINTS")) points_k = atoi(e);
  if (points_k < 1 || points_k > 8 || parts_m <= 0) points_k = 1;
  c->points_per_thread = (unsigned)points_k;
  const std::string cpn = parts_m > 0 ? "ccp" : "cp", kcn = parts_m > 0 ? "ckc" : "kc";
  std::string ib_args, ib_params;
  for (unsigned b = 0; b <= max_back; b++) {
    ib_args += (b ? ", ib" : "ib") + std::to_string(b);
    ib_params += (b ? ", u32 ib" : "u32 ib") + std::to_string(b);
  }
  c->eval_points_per_cta = 0;
  // B2Z_EVALCHECK_TILED (default 0; parts mode, one point per thread, no tile looping): tap loads go to a TILED copy of the three LDE
  // matrices, [tile of 128 points][column][halo + 128 rows] with the halo = 4 * max_back rows that precede the tile (wrapped to
  // the domain), built by one transpose kernel per eval_check.  Every tap of a thread is then its row pointer plus a COMPILE-TIME
  // offset (column * pitch - 16 * back): ONE LDG with an immediate instead of IADD3 + IMAD.X + LDG -- a third of the program's
  // instructions were tap-address arithmetic -- and the eight row-offset registers per thread disappear.
  int tiles_env = 1;
  if (const char* e = getenv("B2Z_EVALCHECK_TILES")) tiles_env = atoi(e);
  //   MEASURED (B200, round 2, profiles/r2_evalcheck_variants.txt): the kernel itself gets 12 % faster (rv32im-scale, po2 20:
  //   35.9 -> 31.4 ms) but the transpose moves 9.5 GB for it (3.8 ms) and holds 5.2 GB per context, and the block job gets
  //   SLOWER (11.3 -> 8.4 M cycles/s, two segments in flight); on the 1 490-step planning program it is 3x slower (7.2 -> 22.5 ms).
  //   A negative result: default OFF, B2Z_EVALCHECK_TILED=1 turns it on.
  const bool tiled = parts_m > 0 && points_k == 1 && tiles_env <= 1 && getenv("B2Z_EVALCHECK_TILED") && getenv("B2Z_EVALCHECK_TILED")[0] == '1';
  if (tiled)
    o << "// tap load from the tiled LDE copy: `p` = this thread's row pointer + a compile-time offset (folded into the LDG immediate)\n"
         "__device__ __forceinline__ u32 ldt(const char* p) { u32 v; asm volatile(\"ld.global.nc.u32 %0, [%1];\" : \"=r\"(v) : \"l\"(p)); return v; }\n";
  const unsigned tile_halo = 4 * max_back, tile_pitch_bytes = 4 * (128 + tile_halo);
  const size_t tws[3] = {c->group_size[0] ? c->group_size[0] : 1, c->group_size[1] ? c->group_size[1] : 1, c->group_size[2] ? c->group_size[2] : 1};
  const size_t tile_ncol = tws[0] + tws[1] + tws[2];
  c->eval_tiled = false;
  const std::string wib_args = ib_args, wib_params = ib_params;  // the witgen parts keep their row offsets
  // B2Z_EVALCHECK_STAGED=1 selects the shared-memory-staged form (codegen_eval_staged).  Measured on B200 (rv32im-scale program,
  // po2 20): per-use global loads 35.7 ms, staged G = 8 / 6 / 4 groups 41.9 / 41.4 / 44.2 ms; planning circuit 1.84 vs 3.95 ms.
  // ncu of the staged kernel (profiles/r2_prof_evalcheck_staged_ncu.txt): DRAM reads 1.08 GB instead of 16.5 GB, 24 % fewer
  // instructions, L2 throughput < 1 % -- but issue 35 % with `no_instruction` at 17 stalled warps per issue: eight different
  // instruction streams per SM, each shared by only four warps, are bound by instruction-FETCH latency.  Off by default.
  bool staged = parts_m > 0 && points_k == 1 && getenv("B2Z_EVALCHECK_STAGED") && getenv("B2Z_EVALCHECK_STAGED")[0] == '1';
  if (staged) staged = codegen_eval_staged(c, o, parts_m);
  if (!staged) {
  std::ostringstream body;  // straight-line mode: kernel body text
  std::vector<Emitter::Stmt> stmts;
  std::vector<size_t> cut_points;  // statement counts at which a cut is allowed
  Val result;
  {
    Emitter em{c, false, &body, {}, 0};
    em.cp_name = cpn;
    em.kc_name = kcn;
    if (parts_m > 0) em.stmts = &stmts;
    if (tiled) {
      em.tiled = true;
      em.tile_flat0[0] = 0;
      em.tile_flat0[1] = tws[0];
      em.tile_flat0[2] = tws[0] + tws[1];
      em.tile_pitch_bytes = tile_pitch_bytes;
    }
    // A mix variable = its accumulated `tot` plus up to three constraints not folded in yet.  Constraint values are base-field
    // words, so a run of AndEqz steps adds pw[e] * val_e to each of the four components: four such terms are accumulated in
    // 64 bits and reduced once (dot4, 10 instructions per component) instead of four products + four modular adds (28).
    // B2Z_EVALCHECK_DOT=0 restores the term-by-term form.
    struct MixAcc {
      Val tot;
      std::vector<std::pair<uint32_t, Opd>> pend;  // (exponent of poly_mix, base-field value)
    };
    const bool use_dot = !(getenv("B2Z_EVALCHECK_DOT") && getenv("B2Z_EVALCHECK_DOT")[0] == '0');
    std::vector<MixAcc> mixv;
    auto flush = [&](MixAcc& m) {
      if (m.pend.empty()) return;
      for (int k = 0; k < 4; k++) {
        std::string expr;
        std::vector<Opd> ops;
        size_t nterm = 0;
        for (auto& pr : m.pend) {
          if (pr.first == 0) continue;  // alpha^0 = 1: handled below
          expr += (nterm ? ", " : "") + kcn + ".pw[" + std::to_string(4 * pr.first + k) + "], " + pr.second.str();
          ops.push_back(pr.second);
          nterm++;
        }
        Opd d;
        bool have = false;
        if (nterm == 1) {
          d = em.emit_n("fmul(" + expr + ")", ops);
          have = true;
        } else if (nterm >= 2) {
          d = em.emit_n("dot" + std::to_string(nterm) + "(" + expr + ")", ops);
          have = true;
        }
        if (have) m.tot.c[k] = em.s_add(m.tot.c[k], d);
      }
      for (auto& pr : m.pend)
        if (pr.first == 0) m.tot.c[0] = em.s_add(m.tot.c[0], pr.second);
      m.pend.clear();
    };
    for (const auto& s : c->steps) {
      switch (s.op) {
        case OP_TRUE: {
          MixAcc z;
          z.tot.ext = true;
          mixv.push_back(z);  // all four components constant 0
        } break;
        case OP_AND_EQZ: {
          Val v = em.use(s.b);
          MixAcc n = mixv[s.a];
          if (use_dot && !v.ext && !v.c[0].is_const) {
            n.pend.push_back({c->mix_exp[s.a], v.c[0]});
            if (n.pend.size() == 4) flush(n);
          } else {
            flush(n);
            Val term = em.v_mul(pw_val(c->mix_exp[s.a], kcn), v);
            n.tot = em.v_add(n.tot, term, false);
          }
          mixv.push_back(n);
          cut_points.push_back(stmts.size());
        } break;
        case OP_AND_COND: {
          MixAcc chain = mixv[s.a], in = mixv[s.c];
          flush(chain);
          flush(in);
          Val inner = em.v_mul(in.tot, em.use(s.b));
          Val term = em.v_mul(inner, pw_val(c->mix_exp[s.a], kcn));
          chain.tot = em.v_add(chain.tot, term, false);
          mixv.push_back(chain);
          cut_points.push_back(stmts.size());
        } break;
        default: em.step(s); break;
      }
    }
    {
      MixAcc r = mixv[c->ret];
      flush(r);
      result = r.tot;
    }
  }
  // parts mode reads its tables from module __constant__ memory (ccp / ckc): the two structs must NOT also travel as kernel
  // parameters (a program of the rv32im size class has ~45 KB of poly_mix powers; the parameter space is 32 KB)
  if (tiled) {
    ib_args = "pt";
    ib_params = "const char* pt";
  }
  const std::string kernel_head_params =
      tiled       ? std::string("u32* __restrict__ check, const char* __restrict__ tiled, u32 iz0, u32 iz1, u32 iz2, u32 iz3, u64 mask")
      : parts_m > 0 ? std::string("u32* __restrict__ check, u32 iz0, u32 iz1, u32 iz2, u32 iz3, u64 mask")
                  : std::string("u32* __restrict__ check, const __grid_constant__ ColPtrs cp, const __grid_constant__ Consts kc, "
                                "u32 iz0, u32 iz1, u32 iz2, u32 iz3, u64 mask");
  // B2Z_EVALCHECK_THREADS = T (parts mode, one point per thread): CTA size.  With T > 128 the CTA's warps are held together by a
  // barrier after every part, so a part's instructions are fetched from L2 once per CTA and stay in the SM's instruction
  // cache while all of its warps run them -- a straight-line program of megabytes otherwise streams once per WARP.
  // Measured on B200 (rv32im-scale program, 27 542 steps, po2 20): T = 128 (no barrier) 52.9 ms, 256: 44.4, 512: 44.6, 1024: 46.6.
  int cta_threads = 256;
  if (const char* e = getenv("B2Z_EVALCHECK_THREADS")) cta_threads = atoi(e);
  if (cta_threads != 128 && cta_threads != 512 && cta_threads != 1024) cta_threads = 256;
  if (parts_m <= 0 || points_k != 1) cta_threads = 128;
  c->eval_threads = (unsigned)cta_threads;
  const bool cta_sync = cta_threads > 128;
  auto emit_tail = [&](std::ostringstream& out) {
    out << "  const u32 iz = (i & 2) ? ((i & 1) ? iz3 : iz2) : ((i & 1) ? iz1 : iz0);\n";
    for (int k = 0; k < 4; k++) {
      std::string v = (result.ext || k == 0) ? result.c[k].str() : std::string("0u");
      out << "  check[" << k << " * domain + i] = fmul(" << v << ", iz);\n";
    }
    out << "}\n\n";
  };
  if (parts_m <= 0) {
    o << "extern \"C\" __global__ void __launch_bounds__(128, " << min_blocks << ") b2z_eval_check(" << kernel_head_params << ") {\n";
    o << "  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;\n  if (i > mask) return;\n  const u64 domain = mask + 1;\n";
    for (unsigned b = 0; b <= max_back; b++) o << "  const u32 ib" << b << " = (u32)((i - " << 4 * b << "ull) & mask) << 2;\n";
    o << body.str();
    emit_tail(o);
  } else {
    o << "__constant__ ColPtrs ccp;\n__constant__ Consts ckc;\n";
    // choose cuts: first allowed cut point at or after every multiple of parts_m statements
    std::vector<size_t> bounds{0};
    for (size_t cp_i : cut_points)
      if (cp_i >= bounds.back() + (size_t)parts_m) bounds.push_back(cp_i);
    if (bounds.back() != stmts.size()) bounds.push_back(stmts.size());
    // definition part of every temporary, and the last part that reads it
    std::map<unsigned, size_t> def_part, last_use_part;
    for (size_t p = 0; p + 1 < bounds.size(); p++)
      for (size_t k = bounds[p]; k < bounds[p + 1]; k++) {
        def_part[stmts[k].id] = p;
        for (unsigned u : stmts[k].uses) last_use_part[u] = p;
      }
    const size_t FINAL = bounds.size();  // pseudo-part: the kernel tail
    for (int k = 0; k < 4; k++)
      if (result.c[k].tid >= 0) last_use_part[(unsigned)result.c[k].tid] = FINAL;
    std::ostringstream main_body;
    std::vector<std::vector<unsigned>> part_ins, part_outs;
    for (size_t p = 0; p + 1 < bounds.size(); p++) {
      std::vector<unsigned> ins, outs;
      std::map<unsigned, bool> seen;
      for (size_t k = bounds[p]; k < bounds[p + 1]; k++) {
        for (unsigned u : stmts[k].uses)
          if (def_part[u] < p && !seen[u]) {
            seen[u] = true;
            ins.push_back(u);
          }
        if (last_use_part.count(stmts[k].id) && last_use_part[stmts[k].id] > p) outs.push_back(stmts[k].id);
      }
      part_ins.push_back(ins);
      part_outs.push_back(outs);
      o << "struct PR" << p << " { u32 v[" << (outs.empty() ? 1 : outs.size()) << "]; };\n";
      o << "static __device__ __noinline__ PR" << p << " part" << p << "(" << ib_params;
      for (unsigned u : ins) o << ", u32 t" << u;
      o << ") {\n";
      for (size_t k = bounds[p]; k < bounds[p + 1]; k++) o << "  const u32 t" << stmts[k].id << " = " << stmts[k].expr << ";\n";
      o << "  PR" << p << " r;\n";
      if (outs.empty()) o << "  r.v[0] = 0u;\n";
      for (size_t j = 0; j < outs.size(); j++) o << "  r.v[" << j << "] = t" << outs[j] << ";\n";
      o << "  return r;\n}\n";
      main_body << "  const PR" << p << " r" << p << " = part" << p << "(" << ib_args;
      for (unsigned u : ins) main_body << ", t" << u;
      main_body << ");\n";
      for (size_t j = 0; j < outs.size(); j++) main_body << "  const u32 t" << outs[j] << " = r" << p << ".v[" << j << "];\n";
    }
    // B2Z_EVALCHECK_TILES = K > 1 (one point per thread, synchronised CTAs): a CTA owns K consecutive tiles of T points and runs
    // EVERY part over all K tiles before it moves on, so a part's instructions are fetched once per K * T points and execute out
    // of the instruction cache K - 1 times out of K (ncu on the rv32im-scale program with K = 1: `no_instruction` is the top
    // stall, issue 55 %).  The few values that are live across parts (the constraint accumulators, pending dot terms) live in
    // shared memory, [slot][tile][thread], instead of registers -- K points per thread at no register cost (round 1's
    // K-points-per-thread emission lost exactly because of the registers).
    int tiles_k = 1;  // measured: K = 1 35.8 ms, K = 2 42.2, K = 3 42.2, K = 4 49.4 (rv32im-scale, po2 20) -- a negative result, kept as a knob
    if (const char* e = getenv("B2Z_EVALCHECK_TILES")) tiles_k = atoi(e);
    if (tiles_k < 1 || tiles_k > 16 || !cta_sync || points_k != 1) tiles_k = 1;
    c->eval_tiles = (unsigned)tiles_k;
    c->eval_smem_bytes = 0;
    if (tiles_k > 1) {
      // slot allocation by liveness: a temporary occupies a slot from its defining part to its last reading part
      const size_t np = bounds.size() - 1;
      std::map<unsigned, unsigned> slot_of;
      std::vector<size_t> free_at;  // per slot: the first part in which it may be reused
      for (size_t p = 0; p < np; p++)
        for (unsigned tid : part_outs[p]) {
          const size_t last = last_use_part[tid];  // a later part, or FINAL
          unsigned sidx = (unsigned)free_at.size();
          for (unsigned q = 0; q < free_at.size(); q++)
            if (free_at[q] <= p) {
              sidx = q;
              break;
            }
          if (sidx == free_at.size()) free_at.push_back(0);
          free_at[sidx] = last == FINAL ? (size_t)-1 : last + 1;  // readable through its last use
          slot_of[tid] = sidx;
        }
      const size_t n_slots = free_at.size() ? free_at.size() : 1;
      c->eval_smem_bytes = 4 * n_slots * (size_t)tiles_k * (size_t)cta_threads;
      const int mb = std::max(1, min_blocks * 128 / cta_threads);
      auto slot = [&](unsigned tid) { return "st[(" + std::to_string(slot_of[tid]) + " * " + std::to_string(tiles_k) + " + k) * " + std::to_string(cta_threads) + " + threadIdx.x]"; };
      o << "// " << n_slots << " shared-memory slots hold the values that cross part boundaries\n";
      o << "extern \"C\" __global__ void __launch_bounds__(" << cta_threads << ", " << mb << ") b2z_eval_check(" << kernel_head_params << ") {\n";
      o << "  extern __shared__ u32 st[];\n  const u64 domain = mask + 1;\n";
      o << "  const u64 base = (u64)blockIdx.x * " << tiles_k * cta_threads << "ull + threadIdx.x;\n";
      auto emit_ib = [&](std::ostringstream& out) {
        out << "    const u64 i0 = base + (u64)k * " << cta_threads << "ull;\n    const u64 i = i0 > mask ? mask : i0;\n";
        for (unsigned b = 0; b <= max_back; b++) out << "    const u32 ib" << b << " = (u32)((i - " << 4 * b << "ull) & mask) << 2;\n";
      };
      for (size_t p = 0; p < np; p++) {
        o << "  __syncthreads();\n#pragma unroll 1\n  for (int k = 0; k < " << tiles_k << "; k++) {\n";
        emit_ib(o);
        o << "    const PR" << p << " r = part" << p << "(" << ib_args;
        for (unsigned u : part_ins[p]) o << ", " << slot(u);
        o << ");\n";
        for (size_t j = 0; j < part_outs[p].size(); j++) o << "    " << slot(part_outs[p][j]) << " = r.v[" << j << "];\n";
        o << "  }\n";
      }
      o << "#pragma unroll 1\n  for (int k = 0; k < " << tiles_k << "; k++) {\n";
      o << "    const u64 i = base + (u64)k * " << cta_threads << "ull;\n    if (i > mask) continue;\n";
      o << "    const u32 iz = (i & 2) ? ((i & 1) ? iz3 : iz2) : ((i & 1) ? iz1 : iz0);\n";
      for (int j = 0; j < 4; j++) {
        std::string v = "0u";
        if (result.ext || j == 0) v = result.c[j].tid >= 0 ? slot((unsigned)result.c[j].tid) : result.c[j].str();
        o << "    check[" << j << " * domain + i] = fmul(" << v << ", iz);\n";
      }
      o << "  }\n}\n\n";
    } else if (points_k == 1) {
      // same register cap as the 128-thread kernel (56): min_blocks CTAs of 128 threads == min_blocks * 128 / T CTAs of T threads
      const int mb = std::max(1, min_blocks * 128 / cta_threads);
      o << "extern \"C\" __global__ void __launch_bounds__(" << cta_threads << ", " << mb << ") b2z_eval_check(" << kernel_head_params << ") {\n";
      if (!cta_sync) {
        o << "  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;\n  if (i > mask) return;\n  const u64 domain = mask + 1;\n";
      } else {
        // every thread must reach the barriers: out-of-range threads (domains smaller than a CTA) evaluate point `mask` and skip the store
        o << "  const u64 i0 = (u64)blockIdx.x * blockDim.x + threadIdx.x;\n  const u64 i = i0 > mask ? mask : i0;\n  const u64 domain = mask + 1;\n";
      }
      if (tiled) {
        o << "  const char* pt = tiled + (i >> 7) * " << (unsigned long long)tile_ncol * tile_pitch_bytes << "ull + ((" << tile_halo
          << "u + (u32)(i & 127)) << 2);\n";
        c->eval_tiled = true;
        c->eval_tile_halo = tile_halo;
        c->eval_tile_ncol = (unsigned)tile_ncol;
      } else {
        for (unsigned b = 0; b <= max_back; b++) o << "  const u32 ib" << b << " = (u32)((i - " << 4 * b << "ull) & mask) << 2;\n";
      }
      if (cta_sync) {
        std::string body = main_body.str(), synced;
        size_t pos = 0, nl;
        while ((nl = body.find('\n', pos)) != std::string::npos) {  // a barrier before every part call
          std::string line = body.substr(pos, nl - pos + 1);
          if (line.find(" = part") != std::string::npos) synced += "  __syncthreads();\n";
          synced += line;
          pos = nl + 1;
        }
        o << synced;
        o << "  if (i0 > mask) return;\n";
      } else {
        o << main_body.str();
      }
      emit_tail(o);
    } else {
      // K points per thread: point k of a thread is (blockIdx.x * K + k) * 128 + threadIdx.x (every k-slice is coalesced);
      // names get a _<k> suffix, part p is called for k = 0..K-1 before part p + 1 starts
      const int reg_blocks = std::max(1, 65536 / (128 * (48 + 12 * points_k)));
      o << "extern \"C\" __global__ void __launch_bounds__(128, " << std::min(min_blocks, reg_blocks) << ") b2z_eval_check("
        << kernel_head_params << ") {\n";
      o << "  const u64 domain = mask + 1;\n";
      for (int k = 0; k < points_k; k++) {
        o << "  const u64 i_" << k << " = (((u64)blockIdx.x * " << points_k << " + " << k << ") * blockDim.x + threadIdx.x);\n";
        for (unsigned b = 0; b <= max_back; b++)
          o << "  const u32 ib" << b << "_" << k << " = (u32)((i_" << k << " - " << 4 * b << "ull) & mask) << 2;\n";
      }
      for (size_t p = 0; p + 1 < bounds.size(); p++) {
        // recompute this part's ins / outs (same rule as above)
        std::vector<unsigned> ins, outs;
        std::map<unsigned, bool> seen;
        for (size_t q = bounds[p]; q < bounds[p + 1]; q++) {
          for (unsigned u : stmts[q].uses)
            if (def_part[u] < p && !seen[u]) {
              seen[u] = true;
              ins.push_back(u);
            }
          if (last_use_part.count(stmts[q].id) && last_use_part[stmts[q].id] > p) outs.push_back(stmts[q].id);
        }
        for (int k = 0; k < points_k; k++) {
          o << "  const PR" << p << " r" << p << "_" << k << " = part" << p << "(";
          for (unsigned b = 0; b <= max_back; b++) o << (b ? ", ib" : "ib") << b << "_" << k;
          for (unsigned u : ins) o << ", t" << u << "_" << k;
          o << ");\n";
          for (size_t j = 0; j < outs.size(); j++) o << "  const u32 t" << outs[j] << "_" << k << " = r" << p << "_" << k << ".v[" << j << "];\n";
        }
      }
      for (int k = 0; k < points_k; k++) {
        o << "  if (i_" << k << " <= mask) {\n";
        o << "    const u32 iz = (i_" << k << " & 2) ? ((i_" << k << " & 1) ? iz3 : iz2) : ((i_" << k << " & 1) ? iz1 : iz0);\n";
        for (int j = 0; j < 4; j++) {
          std::string v = "0u";
          if (result.ext || j == 0) v = result.c[j].tid >= 0 ? "t" + std::to_string(result.c[j].tid) + "_" + std::to_string(k) : result.c[j].str();
          o << "    check[" << j << " * domain + i_" << k << "] = fmul(" << v << ", iz);\n";
        }
        o << "  }\n";
      }
      o << "}\n\n";
    }
  }
  }  // !staged
  // ---- witgen: one thread per active trace row; taps read the trace domain (stride 1) ----
  if (parts_m <= 0) {
    o << "extern \"C\" __global__ void __launch_bounds__(256) b2z_witgen(const __grid_constant__ ColPtrs cp, "
         "const u32* __restrict__ globals, u64 mask, u64 active_rows) {\n";
    o << "  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;\n  if (i >= active_rows) return;\n";
    for (unsigned b = 0; b <= max_back; b++) o << "  const u32 ib" << b << " = (u32)((i - " << b << "ull) & mask) << 2;\n";
    Emitter em{c, true, &o, {}, 0};
    for (const auto& s : c->wit_steps) em.step(s);
    for (size_t k = 0; k < c->wit_store_col.size(); k++) {
      Val v = em.use(c->wit_store_var[k]);
      o << "  const_cast<u32*>(cp.g2[" << c->wit_store_col[k] << "])[i] = " << v.c[0].str() << ";\n";
    }
    o << "}\n";
  } else {
    // Partitioned like eval_check: one huge basic block (a mux circuit's witgen program has tens of thousands of
    // statements) makes ptxas scheduling super-linear -- minutes to hours of compile time.  Cuts are allowed right after a
    // stored value has been computed; column pointers and the globals pointer live in module __constant__ memory.
    o << "__constant__ ColPtrs wcp;\n__constant__ const u32* wglobals;\n";
    std::vector<Emitter::Stmt> wst;
    Emitter em{c, true, &o, {}, 0};
    em.stmts = &wst;
    em.cp_name = "wcp";
    em.glob_name = "wglobals";
    for (const auto& s : c->wit_steps) em.step(s);
    struct Store {
      uint32_t col;
      Opd v;
    };
    std::vector<Store> stores;
    for (size_t k = 0; k < c->wit_store_col.size(); k++) stores.push_back({c->wit_store_col[k], em.use(c->wit_store_var[k]).c[0]});
    std::map<unsigned, size_t> pos_of;  // temporary id -> statement index
    for (size_t k = 0; k < wst.size(); k++) pos_of[wst[k].id] = k;
    std::vector<size_t> cuts;
    for (const auto& st : stores)
      if (st.v.tid >= 0) cuts.push_back(pos_of[(unsigned)st.v.tid] + 1);
    std::sort(cuts.begin(), cuts.end());
    std::vector<size_t> bounds{0};
    for (size_t cp_i : cuts)
      if (cp_i >= bounds.back() + (size_t)parts_m) bounds.push_back(cp_i);
    if (bounds.back() != wst.size()) bounds.push_back(wst.size());
    std::map<unsigned, size_t> def_part, last_use_part;
    for (size_t p = 0; p + 1 < bounds.size(); p++)
      for (size_t k = bounds[p]; k < bounds[p + 1]; k++) {
        def_part[wst[k].id] = p;
        for (unsigned u : wst[k].uses) last_use_part[u] = p;
      }
    const size_t FINAL = bounds.size();
    for (const auto& st : stores)
      if (st.v.tid >= 0) last_use_part[(unsigned)st.v.tid] = FINAL;  // stored by the kernel body right after its part returns
    std::ostringstream wmain;
    for (size_t p = 0; p + 1 < bounds.size(); p++) {
      std::vector<unsigned> ins, outs;
      std::map<unsigned, bool> seen;
      for (size_t k = bounds[p]; k < bounds[p + 1]; k++) {
        for (unsigned u : wst[k].uses)
          if (def_part[u] < p && !seen[u]) {
            seen[u] = true;
            ins.push_back(u);
          }
        if (last_use_part.count(wst[k].id) && last_use_part[wst[k].id] > p) outs.push_back(wst[k].id);
      }
      o << "struct WR" << p << " { u32 v[" << (outs.empty() ? 1 : outs.size()) << "]; };\n";
      o << "static __device__ __noinline__ WR" << p << " wpart" << p << "(" << wib_params;
      for (unsigned u : ins) o << ", u32 t" << u;
      o << ") {\n";
      for (size_t k = bounds[p]; k < bounds[p + 1]; k++) o << "  const u32 t" << wst[k].id << " = " << wst[k].expr << ";\n";
      o << "  WR" << p << " r;\n";
      if (outs.empty()) o << "  r.v[0] = 0u;\n";
      for (size_t jj = 0; jj < outs.size(); jj++) o << "  r.v[" << jj << "] = t" << outs[jj] << ";\n";
      o << "  return r;\n}\n";
      wmain << "  const WR" << p << " w" << p << " = wpart" << p << "(" << wib_args;
      for (unsigned u : ins) wmain << ", t" << u;
      wmain << ");\n";
      for (size_t jj = 0; jj < outs.size(); jj++) wmain << "  const u32 t" << outs[jj] << " = w" << p << ".v[" << jj << "];\n";
      for (const auto& st : stores)
        if (st.v.tid >= 0 && def_part[(unsigned)st.v.tid] == p)
          wmain << "  const_cast<u32*>(wcp.g2[" << st.col << "])[i] = t" << st.v.tid << ";\n";
    }
    for (const auto& st : stores)
      if (st.v.tid < 0) wmain << "  const_cast<u32*>(wcp.g2[" << st.col << "])[i] = " << st.v.str() << ";\n";
    o << "extern \"C\" __global__ void __launch_bounds__(256) b2z_witgen(u64 mask, u64 active_rows) {\n";
    o << "  const u64 i = (u64)blockIdx.x * blockDim.x + threadIdx.x;\n  if (i >= active_rows) return;\n";
    for (unsigned b = 0; b <= max_back; b++) o << "  const u32 ib" << b << " = (u32)((i - " << b << "ull) & mask) << 2;\n";
    o << wmain.str() << "}\n";
  }
  c->source = o.str();
  uint64_t h = 1469598103934665603ull;
  for (unsigned char ch : c->source) {
    h ^= ch;
    h *= 1099511628211ull;
  }
  // the cubin cache key also covers the compiler that would produce the cubin (a cache filled by another NVRTC is not reused)
  int vmaj = 0, vmin = 0;
  if (g_nvrtc.load()) g_nvrtc.Version(&vmaj, &vmin);
  for (unsigned char ch : {(unsigned char)vmaj, (unsigned char)vmin}) {
    h ^= ch;
    h *= 1099511628211ull;
  }
  c->source_hash = h;
}

static std::string cache_dir() {
  if (const char* e = getenv("B2Z_KERNEL_CACHE")) return e;
  Dl_info info;
  if (dladdr((void*)&circuit_codegen, &info) && info.dli_fname) {
    std::string p = info.dli_fname;
    size_t slash = p.rfind('/');
    std::string dir = (slash == std::string::npos ? std::string(".") : p.substr(0, slash)) + "/kernel_cache";
    return dir;
  }
  return "./kernel_cache";
}

b2z_err circuit_cubin(b2z_circuit* c, std::vector<char>& cubin, bool force_recompile) {
  char name[64];
  snprintf(name, sizeof name, "/circuit_%016llx.cubin", (unsigned long long)c->source_hash);
  std::string dir = cache_dir(), path = dir + name;
  if (!force_recompile) {
    std::ifstream f(path, std::ios::binary);
    if (f) {
      cubin.assign(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
      if (!cubin.empty()) return nullptr;
    }
  }
  B2Z_REQUIRE(g_nvrtc.load(), "circuit kernels are not in the cubin cache and libnvrtc.so.12 could not be loaded");
  nvrtcProgram prog = nullptr;
  int rc = g_nvrtc.CreateProgram(&prog, c->source.c_str(), "b2z_circuit.cu", 0, nullptr, nullptr);
  B2Z_REQUIRE(rc == 0, "nvrtcCreateProgram failed");
  const char* opts[] = {"--gpu-architecture=sm_100a", "-lineinfo", "--std=c++17"};
  rc = g_nvrtc.CompileProgram(prog, 3, opts);
  if (rc != 0) {
    size_t ls = 0;
    g_nvrtc.GetProgramLogSize(prog, &ls);
    std::string log(ls, '\0');
    if (ls) g_nvrtc.GetProgramLog(prog, &log[0]);
    g_nvrtc.DestroyProgram(&prog);
    return set_error("NVRTC compile failed: " + log.substr(0, 2000));
  }
  size_t sz = 0;
  g_nvrtc.GetCUBINSize(prog, &sz);
  cubin.resize(sz);
  g_nvrtc.GetCUBIN(prog, cubin.data());
  g_nvrtc.DestroyProgram(&prog);
  mkdir(dir.c_str(), 0755);
  // unique temporary (several ranks may fill a cold cache at once), fsync, then an atomic rename
  std::string tmp = path + ".tmp.XXXXXX";
  int fd = mkstemp(&tmp[0]);
  if (fd >= 0) {
    fchmod(fd, 0644);
    bool ok = write(fd, cubin.data(), cubin.size()) == (ssize_t)cubin.size() && fsync(fd) == 0;
    close(fd);
    if (ok) ok = rename(tmp.c_str(), path.c_str()) == 0;
    if (!ok) unlink(tmp.c_str());
  }
  return nullptr;
}

b2z_err circuit_compile(b2z_ctx* ctx, b2z_circuit* c) {
  std::vector<char> cubin;
  B2Z_TRY(circuit_cubin(c, cubin));
  if (cudaLibraryLoadData(&c->library, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0) != cudaSuccess) {
    // a cached cubin that does not load (partial write of an older build, other toolkit) is thrown away and rebuilt once
    cudaGetLastError();
    c->library = nullptr;
    B2Z_TRY(circuit_cubin(c, cubin, /*force_recompile=*/true));
    B2Z_CUDA(cudaLibraryLoadData(&c->library, cubin.data(), nullptr, nullptr, 0, nullptr, nullptr, 0));
  }
  B2Z_CUDA(cudaLibraryGetKernel(&c->k_eval_check, c->library, "b2z_eval_check"));
  B2Z_CUDA(cudaLibraryGetKernel(&c->k_witgen, c->library, "b2z_witgen"));
  if (c->eval_smem_bytes > 48 * 1024)
    B2Z_CUDA(cudaKernelSetAttributeForDevice(c->k_eval_check, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->eval_smem_bytes, ctx->device));
  if (c->parts_mode) {
    size_t bytes = 0;
    B2Z_CUDA(cudaLibraryGetGlobal(&c->ccp_dev, &bytes, c->library, "ccp"));
    B2Z_CUDA(cudaLibraryGetGlobal(&c->ckc_dev, &bytes, c->library, "ckc"));
    B2Z_CUDA(cudaLibraryGetGlobal(&c->wcp_dev, &bytes, c->library, "wcp"));
    B2Z_CUDA(cudaLibraryGetGlobal(&c->wglobals_dev, &bytes, c->library, "wglobals"));
  }
  c->device = ctx->device;
  if (!c->accum_col.empty()) {
    B2Z_CUDA(cudaMalloc(&c->accum_col_dev, 4 * c->accum_col.size()));
    B2Z_CUDA(cudaMemcpy(c->accum_col_dev, c->accum_col.data(), 4 * c->accum_col.size(), cudaMemcpyHostToDevice));
  }
  if (!c->col_kind.empty()) {
    B2Z_CUDA(cudaMalloc(&c->col_kind_dev, 4 * c->col_kind.size()));
    B2Z_CUDA(cudaMemcpy(c->col_kind_dev, c->col_kind.data(), 4 * c->col_kind.size(), cudaMemcpyHostToDevice));
  }
  return nullptr;
}

}  // namespace b2z
