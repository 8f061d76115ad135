// Per-segment latency distribution of b2z_prove_session (two segments in flight) through the C ABI only -- no Python, no
// torch, no clock sampler: what a C / C++ host of include/b2z.h looks like, and the tool for chasing outlier segments
// (DESIGN.md section 9).  Written at the end of round 2; NOT yet run on a GPU (the round's GPU budget was spent).
//   usage: session_lat <circuit.blob> [po2=20] [segments=32] [in_flight=2]
//   blob:  python -c "from zeth_b200 import synth_circuit as S; import numpy as np; np.asarray(S.circuit_planning().blob, dtype=np.uint32).tofile('planning.blob')"
//   build: g++ -O2 -std=c++17 -o session_lat session_lat.cpp -L../../zeth_b200 -lb2z -Wl,-rpath,'$ORIGIN/../../zeth_b200'
// Every seal is checked with b2z_verify_segment (control root registered from the first witness).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../include/b2z.h"

#define CHK(x)                                     \
  do {                                             \
    b2z_err e_ = (x);                              \
    if (e_) {                                      \
      fprintf(stderr, "%s failed: %s\n", #x, e_);  \
      return 1;                                    \
    }                                              \
  } while (0)

static int run(const std::vector<uint32_t>& blob, unsigned po2, size_t nseg, unsigned in_flight, const char* label) {
  auto t_a = std::chrono::steady_clock::now();
  b2z_session* s = nullptr;
  CHK(b2z_session_create(0, blob.data(), blob.size(), in_flight, &s));
  b2z_circuit* c = b2z_session_circuit(s);
  b2z_ctx* ctx = b2z_session_context(s, 0);
  const size_t n = (size_t)1 << po2, zk = 1994;
  const size_t wa = b2z_circuit_group_size(c, 0), wc = b2z_circuit_group_size(c, 1), wd = b2z_circuit_group_size(c, 2);
  const size_t G = b2z_circuit_globals_size(c);
  const int NW = 4;  // distinct witnesses, used cyclically
  struct W {
    void *code, *data, *glob, *noise;
    std::vector<uint32_t> g;
  } w[NW];
  for (int k = 0; k < NW; k++) {
    CHK(b2z_alloc(ctx, 4 * wc * n, &w[k].code));
    CHK(b2z_alloc(ctx, 4 * wd * n, &w[k].data));
    CHK(b2z_alloc(ctx, 4 * (G ? G : 1), &w[k].glob));
    CHK(b2z_alloc(ctx, 4 * wa * zk, &w[k].noise));
    CHK(b2z_witgen(ctx, c, po2, (uint32_t*)w[k].code, (uint32_t*)w[k].data, (uint32_t*)w[k].glob, (uint32_t*)w[k].noise, 0xB2005000ull + k, zk));
    w[k].g.resize(G ? G : 1);
    CHK(b2z_d2h(ctx, w[k].g.data(), w[k].glob, 4 * (G ? G : 1)));
  }
  uint32_t root[8];
  CHK(b2z_compute_control_root(ctx, po2, (const uint32_t*)w[0].code, wc, root));
  CHK(b2z_session_set_control_root(s, po2, root));
  auto mk = [&](size_t cnt) {
    std::vector<b2z_segment_input> v(cnt);
    for (size_t i = 0; i < cnt; i++) {
      const W& x = w[i % NW];
      v[i] = b2z_segment_input{(uint32_t)i, po2, (const uint32_t*)x.code, (const uint32_t*)x.data, x.g.data(), (const uint32_t*)x.noise, zk, 1};
    }
    return v;
  };
  b2z_receipt* r = nullptr;
  auto warm = mk(8);
  CHK(b2z_prove_session(s, warm.data(), warm.size(), nullptr, 0, &r));
  b2z_receipt_free(r);
  auto t_b = std::chrono::steady_clock::now();
  auto segs = mk(nseg);
  CHK(b2z_prove_session(s, segs.data(), segs.size(), nullptr, 0, &r));
  const double wall = b2z_session_last_wall_ms(s);
  std::vector<double> ms(nseg);
  const size_t cnt = b2z_session_last_segment_ms(s, ms.data(), nseg);
  int bad = 0;
  for (size_t i = 0; i < r->n_segments; i++) bad += b2z_verify_segment(c, r->segments[i].seal, r->segments[i].seal_words) != 0;
  b2z_receipt_free(r);
  std::vector<double> o(ms.begin(), ms.begin() + cnt);
  std::sort(o.begin(), o.end());
  double sum = 0;
  for (double v : o) sum += v;
  printf("%s: po2 %u, %zu segments, %u in flight: wall %.1f ms (%.2f ms/segment, %.2f M cycles/s), latency min %.1f median %.1f p90 %.1f max %.1f, "
         "sum/in_flight %.1f, rejected seals %d, setup %.0f ms\n",
         label, po2, cnt, in_flight, wall, wall / cnt, cnt * (double)n / wall / 1e3, o.front(), o[cnt / 2], o[(cnt * 9) / 10], o.back(), sum / in_flight, bad,
         std::chrono::duration<double, std::milli>(t_b - t_a).count());
  printf("  in order:");
  for (size_t i = 0; i < cnt; i++) printf(" %.0f", ms[i]);
  printf("\n");
  fflush(stdout);
  for (int k = 0; k < NW; k++) {
    b2z_free(ctx, w[k].code);
    b2z_free(ctx, w[k].data);
    b2z_free(ctx, w[k].glob);
    b2z_free(ctx, w[k].noise);
  }
  CHK(b2z_session_destroy(s));
  return bad;
}

int main(int argc, char** argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: %s <circuit.blob> [po2] [segments] [in_flight]\n", argv[0]);
    return 2;
  }
  FILE* f = fopen(argv[1], "rb");
  if (!f) return 2;
  std::vector<uint32_t> blob;
  uint32_t buf[4096];
  size_t got;
  while ((got = fread(buf, 4, 4096, f)) > 0) blob.insert(blob.end(), buf, buf + got);
  fclose(f);
  const unsigned po2 = argc > 2 ? atoi(argv[2]) : 20;
  const size_t nseg = argc > 3 ? atoi(argv[3]) : 32;
  const unsigned inf = argc > 4 ? atoi(argv[4]) : 2;
  return run(blob, po2, nseg, inf, "session");
}
