// STARK segment verifier (host, CPU): the check zeth runs as `receipt.verify(image_id)` per segment
// (/root/reference crates/host/src/bin/cli.rs:103), i.e. risc0-zkp 3.0.2 verify/{mod,merkle,fri,read_iop}.rs
// (scope rows a3 / f2).  Upstream verifies on the CPU as well; this is product
// code (it shares field.cuh / the transcript with the prover) and uses nothing under oracle/.
#include <cstring>

#include "circuit.h"
#include "common.h"

namespace b2z {

namespace {

struct ReadIop {
  const uint32_t* proof;
  size_t len, pos = 0;
  HostRng rng;
  const uint32_t* read(size_t n) {
    if (pos + n > len) return nullptr;
    const uint32_t* p = proof + pos;
    pos += n;
    return p;
  }
};

struct MerkleVerifier {
  size_t rows = 0, cols = 0, top_size = 1;
  std::vector<uint32_t> top;  // 2*top_size digests
  bool init(ReadIop& iop, size_t r, size_t c) {
    rows = r;
    cols = c;
    top_size = (size_t)1 << pp_merkle_top_layer(ilog2(r), B2Z_QUERIES);  // [RECALL med] protocol_params.h
    top.assign(2 * top_size * 8, 0);
    const uint32_t* p = iop.read(8 * top_size);
    if (!p) return false;
    for (size_t i = 0; i < 8 * top_size; i++)
      if (p[i] >= P) return false;  // upstream's hash_pair asserts reduced digest words: x + P would hash like x (malleability)
    memcpy(&top[8 * top_size], p, 32 * top_size);
    for (size_t i = top_size; i-- > 1;) host_hash_pair(&top[8 * 2 * i], &top[8 * (2 * i + 1)], &top[8 * i]);
    iop.rng.mix(&top[8]);
    return true;
  }
  // returns the opened row (inside the seal) or nullptr
  const uint32_t* verify(ReadIop& iop, size_t idx) const {
    if (idx >= rows) return nullptr;
    const uint32_t* row = iop.read(cols);
    if (!row) return nullptr;
    for (size_t i = 0; i < cols; i++)
      if (row[i] >= P) return nullptr;
    uint32_t cur[8], tmp[8];
    host_hash_elem_slice(row, cols, cur);
    idx += rows;
    while (idx >= 2 * top_size) {
      size_t low = idx & 1;
      const uint32_t* other = iop.read(8);
      if (!other) return nullptr;
      for (int i = 0; i < 8; i++)
        if (other[i] >= P) return nullptr;  // non-canonical sibling digest
      idx /= 2;
      if (low) host_hash_pair(other, cur, tmp);
      else host_hash_pair(cur, other, tmp);
      memcpy(cur, tmp, 32);
    }
    return memcmp(&top[8 * idx], cur, 32) == 0 ? row : nullptr;
  }
};

Fp4 poly_eval(const Fp4* coeffs, size_t n, const Fp4& x) {
  Fp4 tot = fp4_zero();
  for (size_t i = n; i-- > 0;) tot = fp4_add(fp4_mul(tot, x), coeffs[i]);
  return tot;
}

// adapter.rs poly_ext: the constraint program on extension-field tap evaluations, MixState{tot, mul} carried explicitly
Fp4 poly_ext(const b2z_circuit* c, const Fp4& poly_mix, const Fp4* eval_u, const uint32_t* globals, const uint32_t* mix) {
  std::vector<Fp4> fpv, tot, mul;
  fpv.reserve(c->steps.size());
  const uint32_t* args[2] = {globals, mix};
  for (const auto& s : c->steps) {
    switch (s.op) {
      case OP_CONST: fpv.push_back(fp4_from_fp(fp_from_u32(s.a))); break;
      case OP_CONST_EXT: fpv.push_back(fp4_make(fp_from_u32(s.a), fp_from_u32(s.b), fp_from_u32(s.c), fp_from_u32(s.d))); break;
      case OP_GET: fpv.push_back(eval_u[s.a]); break;
      case OP_GET_GLOBAL: fpv.push_back(fp4_from_fp(args[s.a][s.b])); break;
      case OP_ADD: fpv.push_back(fp4_add(fpv[s.a], fpv[s.b])); break;
      case OP_SUB: fpv.push_back(fp4_sub(fpv[s.a], fpv[s.b])); break;
      case OP_MUL: fpv.push_back(fp4_mul(fpv[s.a], fpv[s.b])); break;
      case OP_TRUE:
        tot.push_back(fp4_zero());
        mul.push_back(fp4_one());
        break;
      case OP_AND_EQZ:
        tot.push_back(fp4_add(tot[s.a], fp4_mul(mul[s.a], fpv[s.b])));
        mul.push_back(fp4_mul(mul[s.a], poly_mix));
        break;
      case OP_AND_COND:
        tot.push_back(fp4_add(tot[s.a], fp4_mul(fp4_mul(fpv[s.b], tot[s.c]), mul[s.a])));
        mul.push_back(fp4_mul(mul[s.a], mul[s.c]));
        break;
      default: break;
    }
  }
  return tot[c->ret];
}

void interpolate16(Fp4* io) {  // core/ntt.rs interpolate_ntt on 16 ext values, then bit-reverse to natural order
  for (unsigned s = 4; s >= 1; s--) {
    size_t half = (size_t)1 << (s - 1);
    uint32_t step = rou_rev(s);
    for (size_t blk = 0; blk < 16; blk += 2 * half) {
      uint32_t cur = MONT_ONE;
      for (size_t i = 0; i < half; i++) {
        Fp4 a = io[blk + i], b = io[blk + i + half];
        io[blk + i] = fp4_add(a, b);
        io[blk + i + half] = fp4_mul_fp(fp4_sub(a, b), cur);
        cur = fp_mul(cur, step);
      }
    }
  }
  uint32_t norm = fp_inv(fp_from_u32(16));
  for (size_t i = 0; i < 16; i++) io[i] = fp4_mul_fp(io[i], norm);
  for (uint32_t i = 0; i < 16; i++) {
    uint32_t j = bit_rev(i, 4);
    if (i < j) {
      Fp4 t = io[i];
      io[i] = io[j];
      io[j] = t;
    }
  }
}

void hash_info(const uint8_t* info, uint32_t* dg) {
  uint32_t e[16];
  for (int i = 0; i < 16; i++) e[i] = fp_from_u32(info[i]);
  host_hash_elem_slice(e, 16, dg);
}

}  // namespace

// 0 = accepted; otherwise a rejection code (same numbering as the oracle's verifier so tests can compare)
int verify_segment(const b2z_circuit* c, const uint32_t* seal, size_t seal_words) {
  ReadIop iop{seal, seal_words};
  const size_t wa = c->group_size[0], wc = c->group_size[1], wd = c->group_size[2];
  const size_t gsz = c->globals_size, msz = c->mix_size;
  uint32_t dg[8];
  hash_info(reinterpret_cast<const uint8_t*>(PP_PROOF_SYSTEM_INFO), dg);  // [RECALL med] protocol_params.h
  iop.rng.mix(dg);
  hash_info(c->info, dg);
  iop.rng.mix(dg);
  const uint32_t* header = iop.read(gsz + 1);
  if (!header) return 1;
  for (size_t i = 0; i <= gsz; i++)
    if (header[i] >= P) return 2;
  host_hash_elem_slice(header, gsz + 1, dg);
  iop.rng.mix(dg);
  const uint32_t* globals = header;
  uint32_t po2 = fp_to_u32(header[gsz]);
  if (po2 > 24 || po2 < 1) return 3;
  const size_t n = (size_t)1 << po2, domain = 4 * n;
  MerkleVerifier mv[3], mcheck;
  if (!mv[B2Z_GROUP_CODE].init(iop, domain, wc)) return 4;
  {
    // verify/mod.rs check_code(po2, code_merkle.root()): the seal must commit to THIS program's code columns (the
    // control-ID check behind receipt.verify(image_id)).  Fail closed when no control root is registered for the po2.
    if (po2 >= 32 || !c->has_control_root[po2]) return 15;
    if (memcmp(c->control_root[po2], &mv[B2Z_GROUP_CODE].top[8], 32) != 0) return 14;
  }
  if (!mv[B2Z_GROUP_DATA].init(iop, domain, wd)) return 4;
  std::vector<uint32_t> mix(msz + 1);
  for (size_t i = 0; i < msz; i++) mix[i] = iop.rng.random_elem();
  if (!mv[B2Z_GROUP_ACCUM].init(iop, domain, wa)) return 4;
  Fp4 poly_mix = iop.rng.random_ext();
  if (!mcheck.init(iop, domain, B2Z_CHECK_SIZE)) return 4;
  Fp4 z = iop.rng.random_ext();
  uint32_t back_one = rou_rev(po2);
  const size_t n_taps = c->n_taps, n_regs = c->reg_group.size(), n_combos = c->n_combos, tcb = c->tot_combo_backs;
  const uint32_t* cu_words = iop.read(4 * (n_taps + B2Z_CHECK_SIZE));
  if (!cu_words) return 5;
  for (size_t i = 0; i < 4 * (n_taps + B2Z_CHECK_SIZE); i++)
    if (cu_words[i] >= P) return 2;
  const Fp4* coeff_u = reinterpret_cast<const Fp4*>(cu_words);
  host_hash_elem_slice(cu_words, 4 * (n_taps + B2Z_CHECK_SIZE), dg);
  iop.rng.mix(dg);
  std::vector<Fp4> eval_u(n_taps + 1);
  for (size_t r = 0; r < n_regs; r++)
    for (size_t i = 0; i < c->reg_size[r]; i++) {
      Fp4 x = fp4_mul_fp(z, fp_pow(back_one, c->tap_back[c->reg_tap_pos[r] + i]));
      eval_u[c->reg_tap_pos[r] + i] = poly_eval(coeff_u + c->reg_tap_pos[r], c->reg_size[r], x);
    }
  Fp4 result = poly_ext(c, poly_mix, eval_u.data(), globals, mix.data());
  {
    const int* remap = PP_CHECK_REMAP;  // [RECALL med] protocol_params.h
    Fp4 check = fp4_zero();
    for (int i = 0; i < 4; i++) {
      Fp4 zi = fp4_pow(z, (uint64_t)i);
      for (int k = 0; k < 4; k++) {
        Fp4 basis = fp4_zero();
        basis.e[k] = MONT_ONE;
        check = fp4_add(check, fp4_mul(fp4_mul(coeff_u[n_taps + remap[i] + 4 * k], zi), basis));
      }
    }
    Fp4 three_z = fp4_mul_fp(z, MONT_THREE);
    check = fp4_mul(check, fp4_sub(fp4_pow(three_z, n), fp4_one()));
    if (!fp4_eq(check, result)) return 6;
  }
  Fp4 mixv = iop.rng.random_ext();
  std::vector<Fp4> combo_u(tcb + 1, fp4_zero()), tap_mix_pows(n_regs + B2Z_CHECK_SIZE);
  {
    Fp4 cur = fp4_one();
    size_t cur_pos = 0;
    for (size_t r = 0; r < n_regs; r++) {
      for (size_t i = 0; i < c->reg_size[r]; i++) {
        Fp4& d = combo_u[c->combo_begin[c->reg_combo[r]] + i];
        d = fp4_add(d, fp4_mul(cur, coeff_u[cur_pos + i]));
      }
      tap_mix_pows[r] = cur;
      cur = fp4_mul(cur, mixv);
      cur_pos += c->reg_size[r];
    }
    for (size_t i = 0; i < B2Z_CHECK_SIZE; i++) {
      combo_u[tcb] = fp4_add(combo_u[tcb], fp4_mul(cur, coeff_u[cur_pos]));
      cur_pos++;
      tap_mix_pows[n_regs + i] = cur;
      cur = fp4_mul(cur, mixv);
    }
  }
  // fri_verify
  struct Round {
    size_t domain;
    MerkleVerifier merkle;
    Fp4 mix;
  };
  std::vector<Round> rounds;
  size_t degree = n, dom = domain;
  while (degree > B2Z_FRI_MIN_DEGREE) {
    rounds.emplace_back();
    Round& r = rounds.back();
    r.domain = dom / B2Z_FRI_FOLD;
    if (!r.merkle.init(iop, r.domain, B2Z_FRI_FOLD * 4)) return 7;
    r.mix = iop.rng.random_ext();
    dom /= B2Z_FRI_FOLD;
    degree /= B2Z_FRI_FOLD;
  }
  const uint32_t* final_coeffs = iop.read(4 * degree);
  if (!final_coeffs) return 8;
  for (size_t i = 0; i < 4 * degree; i++)
    if (final_coeffs[i] >= P) return 2;
  host_hash_elem_slice(final_coeffs, 4 * degree, dg);
  iop.rng.mix(dg);
  uint32_t gen = rou_fwd(ilog2(dom)), gen0 = rou_fwd(ilog2(domain));
  std::vector<Fp4> poly_buf(degree), tot(n_combos + 1);
  for (size_t i = 0; i < degree; i++)
    for (int j = 0; j < 4; j++) poly_buf[i].e[j] = final_coeffs[j * degree + i];
  const Fp4 z4 = fp4_pow(z, B2Z_INV_RATE);
  for (size_t q = 0; q < B2Z_QUERIES; q++) {
    size_t pos = iop.rng.random_bits(ilog2(domain));
    Fp4 x = fp4_from_fp(fp_pow(gen0, pos));
    const uint32_t* rows[3];
    for (int k = 0; k < 3; k++) {  // [RECALL med] opening order: protocol_params.h PP_QUERY_ORDER
      const int g = PP_QUERY_ORDER[k];
      rows[g] = mv[g].verify(iop, pos);
      if (!rows[g]) return 9;
    }
    const uint32_t* check_row = mcheck.verify(iop, pos);
    if (!check_row) return 9;
    for (auto& t : tot) t = fp4_zero();
    for (size_t r = 0; r < n_regs; r++)
      tot[c->reg_combo[r]] = fp4_add(tot[c->reg_combo[r]], fp4_mul_fp(tap_mix_pows[r], rows[c->reg_group[r]][c->reg_offset[r]]));
    for (size_t i = 0; i < B2Z_CHECK_SIZE; i++) tot[n_combos] = fp4_add(tot[n_combos], fp4_mul_fp(tap_mix_pows[n_regs + i], check_row[i]));
    Fp4 goal = fp4_zero();
    for (size_t k = 0; k < n_combos; k++) {
      size_t b0 = c->combo_begin[k], b1 = c->combo_begin[k + 1];
      Fp4 num = fp4_sub(tot[k], poly_eval(&combo_u[b0], b1 - b0, x));
      Fp4 div = fp4_one();
      for (size_t j = b0; j < b1; j++) div = fp4_mul(div, fp4_sub(x, fp4_mul_fp(z, fp_pow(back_one, c->combo_backs[j]))));
      goal = fp4_add(goal, fp4_mul(num, fp4_inv(div)));
    }
    goal = fp4_add(goal, fp4_mul(fp4_sub(tot[n_combos], combo_u[tcb]), fp4_inv(fp4_sub(x, z4))));
    for (auto& r : rounds) {
      size_t quot = pos / r.domain, group = pos % r.domain;
      const uint32_t* drow = r.merkle.verify(iop, group);
      if (!drow) return 10;
      Fp4 data_ext[16];
      for (int i = 0; i < 16; i++)
        for (int j = 0; j < 4; j++) data_ext[i].e[j] = drow[j * 16 + i];
      if (!fp4_eq(data_ext[quot], goal)) return 11;
      uint32_t inv_wk = fp_pow(rou_rev(ilog2(B2Z_FRI_FOLD * r.domain)), group);
      interpolate16(data_ext);
      goal = poly_eval(data_ext, 16, fp4_mul_fp(r.mix, inv_wk));
      pos = group;
    }
    Fp4 fx = poly_eval(poly_buf.data(), degree, fp4_from_fp(fp_pow(gen, pos)));
    if (!fp4_eq(fx, goal)) return 12;
  }
  if (iop.pos != iop.len) return 13;  // verify_complete
  return 0;
}

}  // namespace b2z
