// Circuit description on the product side: tap set, constraint program, generated kernels.
#pragma once
#include <string>
#include <vector>

#include "common.h"

struct b2z_circuit {
  struct Step {
    uint32_t op, a, b, c, d;
  };
  size_t group_size[3] = {0, 0, 0};
  size_t globals_size = 0, mix_size = 0, n_taps = 0, ret = 0;
  uint8_t info[16] = {};
  std::vector<uint32_t> tap_group, tap_offset, tap_back;
  // registers (runs of taps with equal (group, offset)) and combos (distinct back-lists, lexicographic order)
  std::vector<uint32_t> reg_group, reg_offset, reg_combo, reg_tap_pos, reg_size;
  std::vector<uint32_t> combo_begin, combo_backs;
  size_t n_combos = 0, tot_combo_backs = 0;
  std::vector<Step> steps, wit_steps;
  std::vector<uint32_t> wit_store_col, wit_store_var, accum_col, pin_col, pin_global;
  std::vector<uint32_t> col_kind;  // 0 free, 1 free bit, 2 derived, 3 | idx << 8 | size << 16 | group << 24 one-hot member
  // static analysis
  std::vector<uint8_t> fp_is_ext;
  std::vector<uint32_t> mix_exp;
  size_t max_exp = 0;
  // generated code + loaded module
  std::string source;
  uint64_t source_hash = 0;
  cudaLibrary_t library = nullptr;
  cudaKernel_t k_eval_check = nullptr, k_witgen = nullptr;
  bool parts_mode = false;                    // eval_check split into __noinline__ parts; tables in module constants
  unsigned eval_threads = 128;                // CTA size of the generated eval_check (B2Z_EVALCHECK_THREADS)
  unsigned eval_tiles = 1;                    // tiles of eval_threads points per CTA (B2Z_EVALCHECK_TILES); state in shared memory
  size_t eval_smem_bytes = 0;
  unsigned eval_points_per_cta = 0;
  bool eval_tiled = false;                    // eval_check reads a tiled copy of the LDE matrices (built per launch)
  unsigned eval_tile_halo = 0, eval_tile_ncol = 0;           // staged eval_check: coset points per CTA (0: threads * points * tiles)
  unsigned points_per_thread = 1;             // coset points per thread of the generated eval_check (parts mode, B2Z_EVALCHECK_POINTS)
  void *ccp_dev = nullptr, *ckc_dev = nullptr;  // device addresses of the module's __constant__ tables (parts mode)
  void *wcp_dev = nullptr, *wglobals_dev = nullptr;  // witgen: trace column pointers and the globals pointer (parts mode)
  // control roots (risc0: control IDs): expected Merkle root of the code group per po2, checked by the verifier
  uint32_t control_root[32][8] = {};
  bool has_control_root[32] = {};
  // device-resident tables
  int device = -1;
  uint32_t* accum_col_dev = nullptr;
  uint32_t* col_kind_dev = nullptr;
};

namespace b2z {
enum { OP_CONST = 0, OP_CONST_EXT, OP_GET, OP_GET_GLOBAL, OP_ADD, OP_SUB, OP_MUL, OP_TRUE, OP_AND_EQZ, OP_AND_COND };
b2z_err circuit_parse(b2z_circuit* c, const uint32_t* blob, size_t words);
void circuit_codegen(b2z_circuit* c);
b2z_err circuit_cubin(b2z_circuit* c, std::vector<char>& cubin, bool force_recompile = false);  // cache lookup, else NVRTC
b2z_err circuit_compile(b2z_ctx* ctx, b2z_circuit* c);  // NVRTC (or cubin cache) -> cudaLibrary
size_t circuit_seal_words(const b2z_circuit* c, unsigned po2);
}  // namespace b2z
