// Register-resident integer-pipe microbenchmark for the ALU roofline denominator:
// throughput of IMAD (lo), IMAD.WIDE, IMAD.HI, the Montgomery product and a Shoup constant product on sm_100a.
// Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o intpipe intpipe.cu ; run on the GPU box.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
constexpr uint32_t P = 2013265921u, PINV = 0x88000001u;
__device__ __forceinline__ uint32_t mont(uint32_t a, uint32_t b) {
  uint64_t t = (uint64_t)a * b; uint32_t m = (uint32_t)t * PINV; uint32_t u = __umulhi(m, P); uint32_t r = (uint32_t)(t >> 32) - u;
  return min(r, r + P);
}
__device__ __forceinline__ uint32_t mont2(uint32_t a, uint32_t b) {  // lo + hi as two instructions instead of IMAD.WIDE
  uint32_t lo = a * b, hi = __umulhi(a, b); uint32_t m = lo * PINV; uint32_t u = __umulhi(m, P); uint32_t r = hi - u;
  return min(r, r + P);
}
__device__ __forceinline__ uint32_t shoup(uint32_t x, uint32_t c, uint32_t cp) {  // x*c mod P, cp = floor(c*2^32/P)
  uint32_t q = __umulhi(x, cp); uint32_t r = x * c - q * P; return min(r, r - P);
}
template <int MODE> __global__ void k(uint32_t* out, uint32_t seed, int iters) {
  uint32_t a[8];
  for (int i = 0; i < 8; i++) a[i] = seed * (i + 3) + threadIdx.x;
  uint32_t c = seed | 1, cp = seed * 7 + 1;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) a[i] = a[i] * c + cp;                                   // IMAD lo
      if (MODE == 1) { uint64_t t = (uint64_t)a[i] * c + cp; a[i] = (uint32_t)t ^ (uint32_t)(t >> 32); }  // IMAD.WIDE (+LOP)
      if (MODE == 2) a[i] = __umulhi(a[i], c) + cp;                          // IMAD.HI
      if (MODE == 3) a[i] = mont(a[i], c);
      if (MODE == 4) a[i] = mont2(a[i], c);
      if (MODE == 5) a[i] = shoup(a[i], c, cp);
      if (MODE == 6) { uint32_t r = a[i] + c; a[i] = min(r, r - P); }        // modular add
      if (MODE == 7) a[i] = mont(a[i], a[(i + 1) & 7]);                      // variable x variable
    }
  }
  uint32_t s = 0;
  for (int i = 0; i < 8; i++) s ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char* name, double ops_per_iter) {
  uint32_t* d; cudaMalloc(&d, 4 * 148 * 16 * 256);
  int iters = 4096;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 16, 256>>>(d, 12345, 16);
  cudaEventRecord(e0);
  k<MODE><<<148 * 16, 256>>>(d, 12345, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  double n = 148.0 * 16 * 256 * iters * 8;
  printf("%-28s %8.3f ms  %8.2f Gop/s  (%.2f T instr-equiv/s)\n", name, ms, n / ms / 1e6, n * ops_per_iter / ms / 1e9);
  cudaFree(d);
}
int main() {
  run<0>("IMAD lo", 1); run<1>("IMAD.WIDE(+xor)", 1); run<2>("IMAD.HI", 1); run<3>("mont (wide)", 1); run<4>("mont (lo+hi)", 1);
  run<5>("shoup const", 1); run<6>("modadd", 1); run<7>("mont var*var", 1);
  return 0;
}
