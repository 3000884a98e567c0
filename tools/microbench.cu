// Pipe and modmul throughput microbenchmarks for B200 (sm_100a).
// Evidence for the arithmetic choices in DESIGN.md: prints Gop/s chip-wide for
// raw IMAD / IMAD.WIDE / IADD3 / DFMA chains and G modmul/s for each multiplier
// variant in fp.cuh.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/microbench tools/microbench.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "../ecfft_b200/csrc/fp.cuh"
using namespace ecfft;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

constexpr int ITERS = 512;

__global__ void k_imad(uint32_t* out, uint32_t s) {
  uint32_t a[8];
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x + i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = a[i] * s + a[(i + 1) & 7];
  }
  uint32_t r = 0; for (int i = 0; i < 8; i++) r ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void k_imad_wide(uint64_t* out, uint32_t s) {
  uint64_t a[8];
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x + i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int i = 0; i < 8; i++)
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a[i]) : "r"((uint32_t)a[(i + 1) & 7]), "r"(s));
  }
  uint64_t r = 0; for (int i = 0; i < 8; i++) r ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// the pattern fp.cuh uses: 4-long mad.lo.cc / madc.hi.cc chains (8 independent accumulator windows)
__global__ void k_madcc(uint32_t* out, uint32_t s) {
  uint32_t w[4][9], a[8];
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x * 7 + i;
  for (int c = 0; c < 4; c++) for (int i = 0; i < 9; i++) w[c][i] = c + i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int u = 0; u < 2; u++)
#pragma unroll
      for (int c = 0; c < 4; c++) mad_row4(w[c], a, s + c);
  }
  uint32_t r = 0; for (int c = 0; c < 4; c++) for (int i = 0; i < 9; i++) r ^= w[c][i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void k_iadd3(uint32_t* out, uint32_t s) {
  uint32_t a[8];
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x + i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = a[i] + s + a[(i + 1) & 7];
  }
  uint32_t r = 0; for (int i = 0; i < 8; i++) r ^= a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
__global__ void k_dfma(double* out, double s) {
  double a[8];
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x + i;
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = fma(a[i], s, a[(i + 1) & 7]);
  }
  double r = 0; for (int i = 0; i < 8; i++) r += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}
// IMAD.WIDE and IADD3 issued together (different pipes): does the add ride for free?
__global__ void k_mix(uint64_t* out, uint32_t s) {
  uint64_t a[8]; uint32_t b[8];
  for (int i = 0; i < 8; i++) { a[i] = threadIdx.x + i; b[i] = i * threadIdx.x; }
  for (int it = 0; it < ITERS; it++) {
#pragma unroll
    for (int u = 0; u < 4; u++)
#pragma unroll
      for (int i = 0; i < 8; i++) {
        asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a[i]) : "r"((uint32_t)a[(i + 1) & 7]), "r"(s));
        b[i] = b[i] + s + b[(i + 1) & 7];
      }
  }
  uint64_t r = 0; for (int i = 0; i < 8; i++) r ^= a[i] ^ b[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = r;
}

constexpr int MITERS = 256;
// plain multiply + pseudo-Mersenne reduce, one product per reduction
__global__ void __launch_bounds__(256) k_mul_plain(Fp* out, const Fp* in) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  Fp x = fp_load(in + t), c = fp_load(in + t + 1);
  for (int it = 0; it < MITERS; it++) { x = fp_mul_lazy(x, c); c.v[0] ^= x.v[7]; }
  fp_store(out + t, fp_canon(x));
}
// 2x2 mat-vec with lazy reduction (4 products, 2 reductions): "butterfly"
__global__ void __launch_bounds__(256) k_bfly_lazy(Fp* out, const Fp* in) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  Fp x0 = fp_load(in + t), x1 = fp_load(in + t + 1);
  Fp m00 = fp_load(in + t + 2), m01 = fp_load(in + t + 3), m10 = fp_load(in + t + 4), m11 = fp_load(in + t + 5);
  for (int it = 0; it < MITERS; it++) {
    Fp y0 = fp_dot2_lazy(m00, x0, m01, x1);
    Fp y1 = fp_dot2_lazy(m10, x0, m11, x1);
    x0 = y0; x1 = y1;
  }
  fp_store(out + t, fp_canon(x0)); fp_store(out + t + 1, fp_canon(x1));
}
// normalised butterfly: y0 = x0 + k0*x1, y1 = x0 + k1*x1 (2 products, 2 reductions)
__global__ void __launch_bounds__(256) k_bfly_norm(Fp* out, const Fp* in) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  Fp x0 = fp_load(in + t), x1 = fp_load(in + t + 1);
  Fp k0 = fp_load(in + t + 2), k1 = fp_load(in + t + 3);
  for (int it = 0; it < MITERS; it++) {
    Fp y0 = fp_muladd_lazy(x0, k0, x1);
    Fp y1 = fp_muladd_lazy(x0, k1, x1);
    x0 = y0; x1 = y1;
  }
  fp_store(out + t, fp_canon(x0)); fp_store(out + t + 1, fp_canon(x1));
}
__global__ void __launch_bounds__(256) k_mul_mont(Fp* out, const Fp* in) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  Fp x = fp_load(in + t), c = fp_load(in + t + 1);
  for (int it = 0; it < MITERS; it++) { x = fp_mont_mul(x, c); c.v[0] ^= x.v[7]; }
  fp_store(out + t, x);
}

// ---- the second multiplier: a 256 x 256 -> 520-bit product on the FP64 pipe (52-bit limbs, 5 x 5 partial products,
// each split exactly into high and low 52 bits by two round-toward-zero FMAs; the column sums are integer adds on
// the raw bit patterns).  CORE ONLY: no conversion from / to 32-bit limbs and no reduction mod p, so the rate printed
// is an UPPER bound on what a DFMA-based field product could reach.
__device__ __forceinline__ void dfma_core(const double (&a)[5], const double (&b)[5], long long (&col)[10]) {
  const double C1 = 20282409603651670423947251286016.0;                 // 2^104
  const double C2 = 20282409603651670423947251286016.0 + 4503599627370496.0;  // 2^104 + 2^52
#pragma unroll
  for (int k = 0; k < 10; k++) col[k] = 0;
#pragma unroll
  for (int i = 0; i < 5; i++)
#pragma unroll
    for (int j = 0; j < 5; j++) {
      const double hi = __fma_rz(a[i], b[j], C1);
      const double lo = __fma_rz(a[i], b[j], C2 - hi);
      col[i + j + 1] += __double_as_longlong(hi);
      col[i + j] += __double_as_longlong(lo);
    }
}
__device__ __forceinline__ double limb52(long long c) {   // low 52 bits of a column as an exact double
  return __longlong_as_double((c & 0xFFFFFFFFFFFFFll) | 0x4330000000000000ll) - 4503599627370496.0;
}
__global__ void __launch_bounds__(256) k_mul_dfma(double* out, const double* in) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  double a[5], b[5];
  for (int i = 0; i < 5; i++) { a[i] = limb52((long long)(in[t + i] * 1e9)); b[i] = limb52((long long)(in[t + 5 + i] * 3e9) + i); }
  long long col[10];
  for (int it = 0; it < MITERS; it++) {
    dfma_core(a, b, col);
#pragma unroll
    for (int i = 0; i < 5; i++) a[i] = limb52(col[i] ^ col[i + 5]);
  }
  double r = 0; for (int i = 0; i < 5; i++) r += a[i];
  out[t] = r;
}
// one integer-pipe product and one FP64-pipe product core per iteration, independent chains: is the second one free?
__global__ void __launch_bounds__(256) k_mul_both(Fp* out, const Fp* in, const double* din, int with_dfma) {
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  Fp x = fp_load(in + t), c = fp_load(in + t + 1);
  double a[5], b[5];
  for (int i = 0; i < 5; i++) { a[i] = limb52((long long)(din[t + i] * 1e9)); b[i] = limb52((long long)(din[t + 5 + i] * 3e9) + i); }
  long long col[10];
  for (int it = 0; it < MITERS; it++) {
    x = fp_mul_lazy(x, c); c.v[0] ^= x.v[7];
    if (with_dfma) {
      dfma_core(a, b, col);
#pragma unroll
      for (int i = 0; i < 5; i++) a[i] = limb52(col[i] ^ col[i + 5]);
    }
  }
  double r = 0; for (int i = 0; i < 5; i++) r += a[i];
  x.v[0] ^= (uint32_t)__double_as_longlong(r);
  fp_store(out + t, fp_canon(x));
}

template <typename F>
float time_ms(F launch, int reps = 5) {
  cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  launch(); CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int r = 0; r < reps; r++) {
    CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
  }
  return best;
}

int main() {
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
  int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs, clock %d kHz\n", prop.name, sms, prop.clockRate);
  void* buf; CK(cudaMalloc(&buf, 256u << 20)); CK(cudaMemset(buf, 0x5a, 256u << 20));
  void* inb; CK(cudaMalloc(&inb, 64u << 20)); CK(cudaMemset(inb, 0x37, 64u << 20));
  for (int wps : {4, 8, 16, 32}) {  // warps per SM
    int threads = 256, blocks = sms * wps * 32 / threads * 4;  // 4 waves
    double nthr = (double)threads * blocks;
    double ops = nthr * ITERS * 32.0;
    float ms;
    ms = time_ms([&] { k_imad<<<blocks, threads>>>((uint32_t*)buf, 3); });
    printf("wps=%2d imad        %8.1f Gop/s\n", wps, ops / ms * 1e-6);
    ms = time_ms([&] { k_imad_wide<<<blocks, threads>>>((uint64_t*)buf, 3); });
    printf("wps=%2d imad.wide   %8.1f Gop/s\n", wps, ops / ms * 1e-6);
    ms = time_ms([&] { k_madcc<<<blocks, threads>>>((uint32_t*)buf, 3); });
    printf("wps=%2d madcc(pair) %8.1f Gpair/s\n", wps, nthr * ITERS * 2 * 4 * 4.0 / ms * 1e-6);
    ms = time_ms([&] { k_iadd3<<<blocks, threads>>>((uint32_t*)buf, 3); });
    printf("wps=%2d iadd3       %8.1f Gop/s\n", wps, ops / ms * 1e-6);
    ms = time_ms([&] { k_dfma<<<blocks, threads>>>((double*)buf, 1.000001); });
    printf("wps=%2d dfma        %8.1f Gop/s\n", wps, ops / ms * 1e-6);
    ms = time_ms([&] { k_mix<<<blocks, threads>>>((uint64_t*)buf, 3); });
    printf("wps=%2d wide+iadd3  %8.1f Gpair/s\n", wps, ops / ms * 1e-6);
  }
  for (int bps : {1, 2, 3, 4, 6, 8}) {  // 256-thread blocks per SM resident (grid = 4 waves of that)
    int threads = 256, blocks = sms * bps * 4;
    double nthr = (double)threads * blocks;
    float ms;
    ms = time_ms([&] { k_mul_plain<<<blocks, threads>>>((Fp*)buf, (const Fp*)inb); });
    printf("bps=%d mul_plain   %8.2f Gmul/s\n", bps, nthr * MITERS / ms * 1e-6);
    ms = time_ms([&] { k_mul_mont<<<blocks, threads>>>((Fp*)buf, (const Fp*)inb); });
    printf("bps=%d mul_mont    %8.2f Gmul/s\n", bps, nthr * MITERS / ms * 1e-6);
    ms = time_ms([&] { k_bfly_lazy<<<blocks, threads>>>((Fp*)buf, (const Fp*)inb); });
    printf("bps=%d bfly_lazy   %8.2f Gbfly/s (%.2f Gmul/s)\n", bps, nthr * MITERS / ms * 1e-6, 4 * nthr * MITERS / ms * 1e-6);
    ms = time_ms([&] { k_bfly_norm<<<blocks, threads>>>((Fp*)buf, (const Fp*)inb); });
    printf("bps=%d bfly_norm   %8.2f Gbfly/s (%.2f Gmul/s)\n", bps, nthr * MITERS / ms * 1e-6, 2 * nthr * MITERS / ms * 1e-6);
    ms = time_ms([&] { k_mul_dfma<<<blocks, threads>>>((double*)buf, (const double*)inb); });
    printf("bps=%d mul_dfma_core %6.2f Gmul/s (52-bit limbs, 50 DFMA + 25 DADD + 50 int64 adds; no conversion, no reduction)\n", bps, nthr * MITERS / ms * 1e-6);
    float ms0 = time_ms([&] { k_mul_both<<<blocks, threads>>>((Fp*)buf, (const Fp*)inb, (const double*)inb, 0); });
    float ms1 = time_ms([&] { k_mul_both<<<blocks, threads>>>((Fp*)buf, (const Fp*)inb, (const double*)inb, 1); });
    printf("bps=%d imad only %6.2f Gmul/s; imad + dfma core side by side %6.2f Gmul/s in total (%.2fx the time)\n", bps,
           nthr * MITERS / ms0 * 1e-6, 2 * nthr * MITERS / ms1 * 1e-6, ms1 / ms0);
  }
  return 0;
}
