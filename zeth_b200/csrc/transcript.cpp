// Host side of the Fiat-Shamir transcript: Poseidon2 sponge, hash_pair and Poseidon2Rng on the CPU.
//
// The transcript of one seal needs a few hundred permutations (latency, not throughput), so it runs on the host
// between kernel phases exactly where upstream runs it (risc0-zkp 3.0.2 prove/write_iop.rs,
// core/hash/poseidon2/{mod,rng}.rs; SURVEY.md A.4, scope row a17).  It shares the permutation source with the
// device kernels (poseidon2.cuh) -- this file is product code and does not use anything under oracle/.
#include <cstring>

#include "common.h"
#include "poseidon2.cuh"
#include "poseidon2_consts.h"
#include "poseidon2_f64.cuh"

namespace b2z {

const P2Consts& p2_host_consts() {
  static P2Consts k = [] {
    P2Consts c;
    for (int i = 0; i < P2_NUM_RC; i++) {
      c.rc[i] = fp_from_u32(P2_ROUND_CONSTANTS[i]);
      c.rc_minus_p[i] = (int32_t)(c.rc[i] - P);  // in (-P, 0] as a two's complement value (rc = 0 maps to -P + P... see below)
      if (c.rc[i] == 0) c.rc_minus_p[i] = 0;
    }
    for (int i = 0; i < P2_CELLS; i++) {
      c.diag[i] = fp_from_u32(P2_M_INT_DIAG[i]);
      c.diag_canon[i] = P2_M_INT_DIAG[i];
      c.diag_shoup[i] = (uint32_t)((((unsigned __int128)P2_M_INT_DIAG[i]) << 32) / P);
    }
    return c;
  }();
  return k;
}

void host_poseidon2_mix(uint32_t* s) { p2_permute(s, p2_host_consts()); }

// constants of the FP64 instruction stream (poseidon2_f64.cuh): plain residues as doubles
const P2ConstsF64& p2_host_consts_f64() {
  static P2ConstsF64 k = [] {
    P2ConstsF64 c;
    for (int i = 0; i < P2_NUM_RC; i++) c.rc[i] = (double)P2_ROUND_CONSTANTS[i];
    for (int i = 0; i < P2_CELLS; i++) {
      c.diag[i] = (double)P2_M_INT_DIAG[i];
      c.diag_over_p[i] = (double)P2_M_INT_DIAG[i] / F64_P;
    }
    return c;
  }();
  return k;
}
// the permutation through the FP64 stream, Montgomery words in and out (host copy of what the FP64 warps run)
void host_poseidon2_mix_f64(uint32_t* s) {
  double d[P2_CELLS];
  for (int i = 0; i < P2_CELLS; i++) d[i] = p2f_from_mont(s[i]);
  p2f_permute(d, p2_host_consts_f64());
  for (int i = 0; i < P2_CELLS; i++) s[i] = p2f_to_mont(d[i]);
}

void host_hash_elem_slice(const uint32_t* in, size_t n, uint32_t* digest) {
  uint32_t st[P2_CELLS] = {};
  size_t unmixed = 0;
  for (size_t i = 0; i < n; i++) {
    st[unmixed++] = in[i];
    if (unmixed == P2_RATE) {
      host_poseidon2_mix(st);
      unmixed = 0;
    }
  }
  if (unmixed != 0 || n == 0) {
    for (size_t i = unmixed; i < P2_RATE; i++) st[i] = 0;
    host_poseidon2_mix(st);
  }
  memcpy(digest, st, 32);
}

void host_hash_pair(const uint32_t* a, const uint32_t* b, uint32_t* digest) {
  uint32_t st[P2_CELLS] = {};
  memcpy(st, a, 32);
  memcpy(st + 8, b, 32);
  host_poseidon2_mix(st);
  memcpy(digest, st, 32);
}

void HostRng::mix(const uint32_t* digest) {
  if (pool_used != 0) {
    host_poseidon2_mix(cells);
    pool_used = 0;
  }
  for (int i = 0; i < 8; i++) cells[i] = fp_add(cells[i], digest[i]);
  host_poseidon2_mix(cells);
}
uint32_t HostRng::random_elem() {
  if (pool_used == P2_RATE) {
    host_poseidon2_mix(cells);
    pool_used = 0;
  }
  return cells[pool_used++];
}
Fp4 HostRng::random_ext() {
  Fp4 r;
  for (int i = 0; i < 4; i++) r.e[i] = random_elem();
  return r;
}
uint32_t HostRng::random_bits(unsigned bits) {
  uint32_t val = fp_to_u32(random_elem());
  for (int i = 0; i < PP_RANDOM_BITS_EXTRA_DRAWS; i++) {  // [RECALL lo-med] protocol_params.h
    uint32_t nv = fp_to_u32(random_elem());
    if (val == 0) val = nv;
  }
  return (uint32_t)((((uint64_t)1) << bits) - 1) & val;
}

}  // namespace b2z
