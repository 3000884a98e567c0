// sm_100a EXTEND tile kernel (the dominant kernel of the engine); see DESIGN.md 4.1.
#include <cstdlib>
#include <string>

#include "engine.h"

namespace ecfft {
namespace k {

// ------------------------------------------------------------------------------------------
// EXTEND tile kernel.
//
// A vector of length h is viewed through the levels of the butterfly network: the level with
// half-stride 2^j pairs positions p and p + 2^j (bit j of p clear) and uses matrix
// M[2^(j+1) + 2*(p mod 2^j) + skip] of the chain level's matrix BinaryTree (layer offset =
// block size, reference src/utils.rs:248-252).  A CTA owns every element that agrees on all
// position bits outside [j_lo, j_hi) and on the high column bits: 2^(j_hi-j_lo) rows of
// C = 2^log_c contiguous elements.  It runs the decompose levels j = j_hi-1 .. j_lo, then (for
// the innermost pass) the recombine levels j = j_lo .. j_hi-1, with one __syncthreads() per
// level and no global traffic in between.  Blocks are ordered batch-major so CTAs resident at
// the same time share matrix lines in L2/L1.
// ------------------------------------------------------------------------------------------
struct TileParams {
  const Fp* in;
  Fp* out;
  const Fp* dmat;   // MODE 0: decompose matrices | MODE 1: tw_d[source] ({c, -s0} per butterfly)
  const Fp* rmat;   // MODE 0: recombine matrices | MODE 1: tw_r[target] ({s0, s1} per butterfly)
  const Fp* pre;    // MODE 1: per-position scale applied on load (1/Gamma^source) or null
  const Fp* post;   // MODE 1: per-position scale applied on store (Gamma^target) or null
  unsigned long long nvec;
  unsigned long long total;  // nvec * h, guards the ragged last tile of the packed mode
  uint32_t log_h, j_lo, j_hi, log_c;
  uint32_t log_t;            // tile holds 2^log_t elements
  uint32_t packed;           // 1: h <= tile, a tile is 2^(log_t-log_h) whole consecutive vectors
  uint32_t norm;             // 1: normalised butterflies (MODE 1)
  uint32_t do_d, do_r, skip_d, skip_r;
};

// MODE 0 — the reference's 2x2 mat-vec (src/utils.rs:338-347): 4 products, 2 lazy reductions
__device__ __forceinline__ void butterfly_matrix(Fp* s, uint32_t e_lo, uint32_t e_hi, const Fp* m) {
  Fp m0 = fp_load_ro(m), m1 = fp_load_ro(m + 1);
  Fp x0 = s[e_lo], x1 = s[e_hi];
  Fp y0 = fp_dot2_lazy(m0, x0, m1, x1);
  Fp m2 = fp_load_ro(m + 2), m3 = fp_load_ro(m + 3);
  s[e_lo] = y0;
  Fp y1 = fp_dot2_lazy(m2, x0, m3, x1);
  s[e_hi] = y1;
}
// MODE 1 recombine: [[1, s0], [1, s1]] — the two outputs are x_p + s*x_q at the pair's two nodes
__device__ __forceinline__ void butterfly_norm_r_v(Fp* s, uint32_t e_lo, uint32_t e_hi, const Fp& s0, const Fp& s1) {
  Fp xp = s[e_lo], xq = s[e_hi];
  s[e_lo] = fp_muladd_lazy(xp, s0, xq);
  s[e_hi] = fp_muladd_lazy(xp, s1, xq);
}
__device__ __forceinline__ void butterfly_norm_r(Fp* s, uint32_t e_lo, uint32_t e_hi, const Fp* tw) {
  butterfly_norm_r_v(s, e_lo, e_hi, fp_load_ro(tw), fp_load_ro(tw + 1));
}
// MODE 1 decompose: inverse of [[1, s0], [1, s1]].
#ifndef ECFFT_D_DIFFFORM
// Sum form: input (column) scalings commute backwards through the decompose phase, so with
// x^_p = -c x_p, x^_q = c x_q (c = 1/(s1-s0), folded into the 1/Gamma pre-scale table) the pair is
// y_q = x^_p + x^_q (no multiplication), y_p = -(s1 x^_p + s0 x^_q) (two products, ONE reduction).
// tw = {-s1, -s0}
__device__ __forceinline__ void butterfly_norm_d_v(Fp* s, uint32_t e_lo, uint32_t e_hi, const Fp& ns1, const Fp& ns0) {
  Fp xp = s[e_lo], xq = s[e_hi];
  s[e_hi] = fp_add_lazy(xp, xq);
  s[e_lo] = fp_dot2_lazy(ns1, xp, ns0, xq);
}
__device__ __forceinline__ void butterfly_norm_d(Fp* s, uint32_t e_lo, uint32_t e_hi, const Fp* tw) {
  butterfly_norm_d_v(s, e_lo, e_hi, fp_load_ro(tw), fp_load_ro(tw + 1));
}
#else
// y_q = (x_q - x_p)/(s1 - s0), y_p = x_p - s0*y_q;  tw = {1/(s1-s0), -s0}
__device__ __forceinline__ void butterfly_norm_d_v(Fp* s, uint32_t e_lo, uint32_t e_hi, const Fp& c, const Fp& ns0) {
  Fp xp = fp_canon(s[e_lo]), xq = s[e_hi];
  Fp yq = fp_mul_lazy(c, fp_sub_lazy(xq, xp));
  s[e_hi] = yq;
  s[e_lo] = fp_muladd_lazy(xp, ns0, yq);
}
__device__ __forceinline__ void butterfly_norm_d(Fp* s, uint32_t e_lo, uint32_t e_hi, const Fp* tw) {
  butterfly_norm_d_v(s, e_lo, e_hi, fp_load_ro(tw), fp_load_ro(tw + 1));
}
#endif

template <int MODE, int NT, int MINB, int UNROLL, int PF = 0>
__global__ void __launch_bounds__(NT, MINB) k_extend_tile(TileParams p) {
  static_assert(MODE == 0 || MODE == 1, "butterfly mode");
  extern __shared__ uint4 smem_raw[];
  Fp* s = reinterpret_cast<Fp*>(smem_raw);
  const uint32_t T = 1u << p.log_t;
  const uint32_t C = 1u << p.log_c;
  const unsigned long long hmask = (1ull << p.log_h) - 1;
  unsigned long long pos0, gbase;
  if (p.packed) {  // j_lo = 0, C = 1: element e of the tile is global element gbase + e
    pos0 = 0;
    gbase = (unsigned long long)blockIdx.x << p.log_t;
  } else {
    const unsigned long long v = blockIdx.x % p.nvec;
    const unsigned long long tile = blockIdx.x / p.nvec;
    const uint32_t ncg_log = p.j_lo - p.log_c;
    const unsigned long long cg = tile & ((1ull << ncg_log) - 1);
    const unsigned long long q_hi = tile >> ncg_log;
    pos0 = (q_hi << p.j_hi) + (cg << p.log_c);  // position within the vector of tile element 0
    gbase = (v << p.log_h) + pos0;
  }

  for (uint32_t e = threadIdx.x; e < T; e += NT) {
    uint32_t r = e >> p.log_c, c = e & (C - 1);
    unsigned long long g = gbase + ((unsigned long long)r << p.j_lo) + c;
    Fp x = g < p.total ? fp_load(p.in + g) : fp_zero();
    if (MODE == 1 && p.pre) x = fp_mul_lazy(x, fp_load_ro(p.pre + (g & hmask)));
    s[e] = x;
  }
  __syncthreads();

  if (p.do_d) {
    for (int j = (int)p.j_hi - 1; j >= (int)p.j_lo; j--) {
      const uint32_t sh = (uint32_t)j - p.j_lo + p.log_c;  // bit of the tile index that this level pairs
      const unsigned long long jmask = (1ull << j) - 1;
      const Fp* layer = MODE == 0 ? p.dmat + 4 * ((2ull << j) + p.skip_d) : p.dmat + 2 * (1ull << j);
      if (MODE == 1 && PF) {  // software pipelining: the next pair's twiddles load while this pair computes
        uint32_t b = threadIdx.x;
        uint32_t e_nx = ((b >> sh) << (sh + 1)) | (b & ((1u << sh) - 1));
        const Fp* tw = layer + 2 * ((pos0 + ((unsigned long long)(e_nx >> p.log_c) << p.j_lo) + (e_nx & (C - 1))) & jmask);
        Fp t0 = fp_load_ro(tw), t1 = fp_load_ro(tw + 1);
#pragma unroll 1
        for (; b < T / 2; b += NT) {
          const uint32_t e_lo = e_nx;
          const Fp c0 = t0, c1 = t1;
          if (b + NT < T / 2) {
            const uint32_t bn = b + NT;
            e_nx = ((bn >> sh) << (sh + 1)) | (bn & ((1u << sh) - 1));
            tw = layer + 2 * ((pos0 + ((unsigned long long)(e_nx >> p.log_c) << p.j_lo) + (e_nx & (C - 1))) & jmask);
            t0 = fp_load_ro(tw);
            t1 = fp_load_ro(tw + 1);
          }
          butterfly_norm_d_v(s, e_lo, e_lo + (1u << sh), c0, c1);
        }
      } else {
#pragma unroll UNROLL
      for (uint32_t b = threadIdx.x; b < T / 2; b += NT) {
        uint32_t e_lo = ((b >> sh) << (sh + 1)) | (b & ((1u << sh) - 1));
        uint32_t r = e_lo >> p.log_c, c = e_lo & (C - 1);
        unsigned long long i = (pos0 + ((unsigned long long)r << p.j_lo) + c) & jmask;
        if (MODE == 0)
          butterfly_matrix(s, e_lo, e_lo + (1u << sh), layer + 8 * i);
        else
          butterfly_norm_d(s, e_lo, e_lo + (1u << sh), layer + 2 * i);
      }
      }
      __syncthreads();
    }
  }
  if (p.do_r) {
    for (uint32_t j = p.j_lo; j < p.j_hi; j++) {
      const uint32_t sh = j - p.j_lo + p.log_c;
      const unsigned long long jmask = (1ull << j) - 1;
      const Fp* layer = MODE == 0 ? p.rmat + 4 * ((2ull << j) + p.skip_r) : p.rmat + 2 * (1ull << j);
      if (MODE == 1 && PF) {  // software pipelining: the next pair's twiddles load while this pair computes
        uint32_t b = threadIdx.x;
        uint32_t e_nx = ((b >> sh) << (sh + 1)) | (b & ((1u << sh) - 1));
        const Fp* tw = layer + 2 * ((pos0 + ((unsigned long long)(e_nx >> p.log_c) << p.j_lo) + (e_nx & (C - 1))) & jmask);
        Fp t0 = fp_load_ro(tw), t1 = fp_load_ro(tw + 1);
#pragma unroll 1
        for (; b < T / 2; b += NT) {
          const uint32_t e_lo = e_nx;
          const Fp c0 = t0, c1 = t1;
          if (b + NT < T / 2) {
            const uint32_t bn = b + NT;
            e_nx = ((bn >> sh) << (sh + 1)) | (bn & ((1u << sh) - 1));
            tw = layer + 2 * ((pos0 + ((unsigned long long)(e_nx >> p.log_c) << p.j_lo) + (e_nx & (C - 1))) & jmask);
            t0 = fp_load_ro(tw);
            t1 = fp_load_ro(tw + 1);
          }
          butterfly_norm_r_v(s, e_lo, e_lo + (1u << sh), c0, c1);
        }
      } else {
#pragma unroll UNROLL
      for (uint32_t b = threadIdx.x; b < T / 2; b += NT) {
        uint32_t e_lo = ((b >> sh) << (sh + 1)) | (b & ((1u << sh) - 1));
        uint32_t r = e_lo >> p.log_c, c = e_lo & (C - 1);
        unsigned long long i = (pos0 + ((unsigned long long)r << p.j_lo) + c) & jmask;
        if (MODE == 0)
          butterfly_matrix(s, e_lo, e_lo + (1u << sh), layer + 8 * i);
        else
          butterfly_norm_r(s, e_lo, e_lo + (1u << sh), layer + 2 * i);
      }
      }
      __syncthreads();
    }
  }
  for (uint32_t e = threadIdx.x; e < T; e += NT) {
    uint32_t r = e >> p.log_c, c = e & (C - 1);
    unsigned long long g = gbase + ((unsigned long long)r << p.j_lo) + c;
    if (g < p.total) {
      Fp x = s[e];
      if (MODE == 1 && p.post) x = fp_mul_lazy(x, fp_load_ro(p.post + (g & hmask)));
      fp_store(p.out + g, fp_canon(x));
    }
  }
}

// Launch-shape variants (ECFFT_B200_TILE_VARIANT, measured in profiles/):
//   0: 256 threads, 2048-element tile, butterfly loop unrolled x4 (2 CTAs/SM)
//   1: 256 threads, 2048-element tile, no unrolling (3 CTAs/SM fit)
//   2: as 1 with __launch_bounds__(256, 3)
//   3: 512 threads, 1 CTA/SM, 4096-element tile (12 + 12 levels in the inner pass)
//   4: 512 threads, 1 CTA/SM, 2048-element tile
//   5: as 1 with the butterfly loop unrolled x2
//   6: 128 threads, 1024-element tile, __launch_bounds__(128, 4)
//   7: 128 threads, 1024-element tile, __launch_bounds__(128, 6): 72 registers, 7 CTAs/SM — default, fastest
//   8: as 1 with software-prefetched twiddles
static int tile_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ECFFT_B200_TILE_VARIANT");
    v = e ? atoi(e) : 7;
    if (v < 0 || v > 8) v = 7;
  }
  return v;
}
static uint32_t log_tile() {
  int v = tile_variant();
  return v == 3 ? 12 : (v == 6 || v == 7) ? 10 : 11;
}

template <int MODE, int NT, int MINB, int UNROLL, int PF = 0>
static void launch_variant(const TileParams& p, size_t tiles, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    ECFFT_CUDA(cudaFuncSetAttribute(k_extend_tile<MODE, NT, MINB, UNROLL, PF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((1u << 12) * sizeof(Fp))));
    configured = true;
  }
  k_extend_tile<MODE, NT, MINB, UNROLL, PF><<<(unsigned)tiles, NT, ((size_t)sizeof(Fp)) << p.log_t, st>>>(p);
}

static void launch_tile(const TileParams& p, cudaStream_t st) {
  size_t tiles = (p.total + ((size_t)1 << p.log_t) - 1) >> p.log_t;
  if (tiles > 0x7fffffffull) throw Error(ERR_INVALID_ARG, "extend: grid too large");
  const bool timed = prof::enabled();
  if (timed) {
    // algorithmic bytes: every fused level reads and writes each element once (64 B) and reads its
    // 2^j matrices (128 B each) once
    double levels = (double)(p.j_hi - p.j_lo) * (p.do_d + p.do_r);
    double mats = 0;
    for (uint32_t j = p.j_lo; j < p.j_hi; j++) mats += (double)(p.do_d + p.do_r) * 128.0 * (double)(1ull << j);
    prof::record_begin(prof::EXTEND_TILE, levels * 64.0 * (double)p.total + mats, st);
  }
  if (!p.norm) {
    launch_variant<0, 256, 2, 4>(p, tiles, st);
  } else {
    switch (tile_variant()) {
      case 0: launch_variant<1, 256, 2, 4>(p, tiles, st); break;
      case 1: launch_variant<1, 256, 2, 1>(p, tiles, st); break;
      case 2: launch_variant<1, 256, 3, 1>(p, tiles, st); break;
      case 3: launch_variant<1, 512, 1, 1>(p, tiles, st); break;
      case 4: launch_variant<1, 512, 1, 1>(p, tiles, st); break;
      case 5: launch_variant<1, 256, 2, 2>(p, tiles, st); break;
      case 6: launch_variant<1, 128, 4, 1>(p, tiles, st); break;
      case 8: launch_variant<1, 256, 2, 1, 1>(p, tiles, st); break;
      default: launch_variant<1, 128, 6, 1>(p, tiles, st); break;
    }
  }
  if (timed) prof::record_end(st);
  prof::count_launch();
  ECFFT_CUDA(cudaGetLastError());
}

int butterfly_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("ECFFT_B200_BUTTERFLY");
    mode = (e && std::string(e) == "matrix") ? 0 : 1;
  }
  return mode;
}

// Schedules the tile passes for all levels j < log_h of vectors of length 2^log_h (p holds the tables,
// mode and skips; pre/post are the per-position scales of the first load / last store or null).
static void run_passes(TileParams p, const Fp* in, Fp* out, uint32_t log_h, size_t nvec, const Fp* pre, const Fp* post, cudaStream_t st) {
  const uint32_t LT = log_tile();
  p.pre = nullptr;
  p.post = nullptr;
  p.nvec = nvec;
  p.total = nvec << log_h;
  p.log_h = log_h;
  p.log_t = LT;
  p.packed = 0;
  if (log_h <= LT) {
    // whole vectors fit a tile: pack 2^(log_t-log_h) consecutive vectors per CTA
    p.in = in; p.out = out;
    p.j_lo = 0; p.j_hi = log_h; p.log_c = 0; p.do_d = 1; p.do_r = 1;
    p.packed = 1;
    p.log_t = log_h;
    while (p.log_t < LT && ((size_t)1 << p.log_t) < p.total) p.log_t++;
    p.pre = pre;
    p.post = post;
    launch_tile(p, st);
    return;
  }
  // outer decompose passes (strided tiles), inner fused pass, outer recombine passes
  const uint32_t outer = log_h - LT;
  const uint32_t kmax = LT - 5;                   // rows of 32 contiguous elements (1 KiB)
  const uint32_t npass = (outer + kmax - 1) / kmax;
  std::vector<uint32_t> bounds;                   // j boundaries from log_h down to LT
  bounds.push_back(log_h);
  for (uint32_t i = 1; i <= npass; i++) bounds.push_back(log_h - (outer * i) / npass);
  const Fp* src = in;
  for (uint32_t i = 0; i < npass; i++) {
    p.in = src; p.out = out;
    p.j_hi = bounds[i]; p.j_lo = bounds[i + 1]; p.log_c = LT - (p.j_hi - p.j_lo);
    p.do_d = 1; p.do_r = 0;
    p.pre = i == 0 ? pre : nullptr;  // 1/Gamma^source on the very first load
    launch_tile(p, st);
    src = out;
  }
  p.pre = nullptr;
  p.in = out; p.out = out; p.j_lo = 0; p.j_hi = LT; p.log_c = 0; p.do_d = 1; p.do_r = 1;
  launch_tile(p, st);
  for (uint32_t i = npass; i-- > 0;) {
    p.in = out; p.out = out;
    p.j_hi = bounds[i]; p.j_lo = bounds[i + 1]; p.log_c = LT - (p.j_hi - p.j_lo);
    p.do_d = 0; p.do_r = 1;
    p.post = i == 0 ? post : nullptr;  // Gamma^target on the very last store
    launch_tile(p, st);
  }
}

void extend(const Level& lv, const Fp* in, Fp* out, uint32_t log_h, size_t nvec, Moiety target, cudaStream_t st, bool unscaled_out) {
  if (nvec == 0) return;
  if (log_h == 0) {  // extend_impl n == 1: identity, fftree.rs:74-76
    if (in != out) ECFFT_CUDA(cudaMemcpyAsync(out, in, nvec * sizeof(Fp), cudaMemcpyDeviceToDevice, st));
    return;
  }
  TileParams p;
  const Moiety source = target == S1 ? S0 : S1;
  const bool norm = butterfly_mode() == 1 && lv.tw_r[target] && lv.tw_d[source] && lv.gam[target] && lv.gami[source];
  if (unscaled_out && !norm) throw Error(ERR_INVALID_ARG, "extend: unscaled output needs the normalised tables");
  p.norm = norm ? 1 : 0;
  p.dmat = norm ? lv.tw_d[source] : lv.dmat;
  p.rmat = norm ? lv.tw_r[target] : lv.rmat;
  p.skip_d = target == S0 ? 1 : 0;  // fftree.rs:87-90
  p.skip_r = target == S1 ? 1 : 0;  // fftree.rs:108-111
  run_passes(p, in, out, log_h, nvec, norm ? lv.gami[source] : nullptr, (norm && !unscaled_out) ? lv.gam[target] : nullptr, st);
}

// Multi-GPU building block: the rank-local levels (half-strides < 2^log_len) of the normalised
// EXTEND -> S1 of a longer vector whose contiguous chunk of 2^log_len elements this rank holds.
// Twiddles depend only on the position modulo the half-stride, so the chunk behaves like a vector of
// its own length with the long vector's tables; the diagonal scalings are applied by the caller.
void extend_sub(const Level& lv, const Fp* in, Fp* out, uint32_t log_len, cudaStream_t st) {
  if (!(lv.tw_r[1] && lv.tw_d[0])) throw Error(ERR_MISSING_TABLES, "extend_sub: normalised tables missing");
  if (log_len == 0) {
    if (in != out) ECFFT_CUDA(cudaMemcpyAsync(out, in, sizeof(Fp), cudaMemcpyDeviceToDevice, st));
    return;
  }
  TileParams p;
  p.norm = 1;
  p.dmat = lv.tw_d[0];
  p.rmat = lv.tw_r[1];
  p.skip_d = 0;
  p.skip_r = 1;
  run_passes(p, in, out, log_len, 1, nullptr, nullptr, st);
}

// ------------------------------------------------------------------------------------------
// k_enter_small: every ENTER recursion depth with block size m <= 2^LOG_SMALL in ONE launch.
// A CTA owns 2^log_t consecutive coefficients; per depth it (1) copies the current vectors scaled by
// 1/Gamma^0 into a work buffer, (2) runs the normalised decompose + recombine levels of EXTEND -> S1 on
// it, (3) combines (src/fftree.rs:155-159) into a third buffer — all in shared memory.  Replaces
// 2 launches and ~160 B/element of HBM traffic per depth of the generic path.
// ------------------------------------------------------------------------------------------
static constexpr uint32_t LOG_SMALL = 10;  // 3 buffers x 1024 x 32 B = 96 KiB per CTA, 2 CTAs/SM
struct SmallLevel {
  const Fp* tw_d0;   // level's tw_d[0]
  const Fp* tw_r1;   // tw_r[1]
  const Fp* gami0;
  const Fp* gam1;
  const Fp* gx;
  const Fp* xnn;
};
struct SmallParams {
  const Fp* in;
  Fp* out;
  unsigned long long total;
  uint32_t log_t;           // tile size (<= LOG_SMALL)
  uint32_t lvl_lo, lvl_hi;  // depths with log2(m) in (lvl_lo, lvl_hi]
  SmallLevel lv[LOG_SMALL + 1];  // indexed by log2(m)
};

__global__ void __launch_bounds__(256, 2) k_enter_small(const __grid_constant__ SmallParams p) {
  extern __shared__ uint4 smem_raw[];
  const uint32_t T = 1u << p.log_t;
  Fp* A = reinterpret_cast<Fp*>(smem_raw);
  Fp* W = A + T;
  Fp* B = W + T;
  const unsigned long long gbase = (unsigned long long)blockIdx.x << p.log_t;
  for (uint32_t e = threadIdx.x; e < T; e += 256) A[e] = gbase + e < p.total ? fp_load(p.in + gbase + e) : fp_zero();
  __syncthreads();
  for (uint32_t lm = p.lvl_lo + 1; lm <= p.lvl_hi; lm++) {
    const SmallLevel& lv = p.lv[lm];
    const uint32_t L = lm - 1, h = 1u << L;
    if (L > 0) {
      for (uint32_t e = threadIdx.x; e < T; e += 256) W[e] = fp_mul_lazy(A[e], fp_load_ro(lv.gami0 + (e & (h - 1))));
      __syncthreads();
      for (int j = (int)L - 1; j >= 0; j--) {
        const Fp* layer = lv.tw_d0 + 2 * (1u << j);
#pragma unroll 1
        for (uint32_t b = threadIdx.x; b < T / 2; b += 256) {
          uint32_t e_lo = ((b >> j) << (j + 1)) | (b & ((1u << j) - 1));
          butterfly_norm_d(W, e_lo, e_lo + (1u << j), layer + 2 * (e_lo & ((1u << j) - 1)));
        }
        __syncthreads();
      }
      for (uint32_t j = 0; j < L; j++) {
        const Fp* layer = lv.tw_r1 + 2 * (1u << j);
#pragma unroll 1
        for (uint32_t b = threadIdx.x; b < T / 2; b += 256) {
          uint32_t e_lo = ((b >> j) << (j + 1)) | (b & ((1u << j) - 1));
          butterfly_norm_r(W, e_lo, e_lo + (1u << j), layer + 2 * (e_lo & ((1u << j) - 1)));
        }
        __syncthreads();
      }
    }
    const Fp* Wsrc = L > 0 ? W : A;  // EXTEND of a length-1 vector is the identity (fftree.rs:74-76)
#pragma unroll 1
    for (uint32_t idx = threadIdx.x; idx < T / 2; idx += 256) {
      uint32_t blk = idx >> L, i = idx & (h - 1), off = blk << lm;
      Fp u0 = A[off + i], v0 = A[off + h + i];
      B[off + 2 * i] = fp_muladd_lazy(u0, v0, fp_load_ro(lv.xnn + 2 * i));
      Fp u1 = Wsrc[off + i], v1 = Wsrc[off + h + i];
      B[off + 2 * i + 1] = fp_dot2_lazy(fp_load_ro(lv.gam1 + i), u1, fp_load_ro(lv.gx + i), v1);
    }
    __syncthreads();
    Fp* sw = A; A = B; B = sw;
  }
  for (uint32_t e = threadIdx.x; e < T; e += 256)
    if (gbase + e < p.total) fp_store(p.out + gbase + e, fp_canon(A[e]));
}

// runs depths m_lo < m <= m_hi (m_hi <= 2^LOG_SMALL, m_hi <= n) of the bottom-up ENTER; levels[k] is the
// chain level with 2^k leaves.  Returns false when the normalised tables are not available.
bool enter_small(const Level* levels, const Fp* in, Fp* out, size_t n, size_t m_lo, size_t m_hi, cudaStream_t st) {
  if (butterfly_mode() != 1) return false;
  uint32_t lo = 0, hi = 0;
  while (((size_t)1 << lo) < m_lo) lo++;
  while (((size_t)1 << hi) < m_hi) hi++;
  if (hi > LOG_SMALL || hi <= lo) return false;
  SmallParams p;
  p.in = in;
  p.out = out;
  p.total = n;
  p.log_t = hi;  // a tile must hold whole blocks of the largest depth
  while (p.log_t < LOG_SMALL && ((size_t)1 << p.log_t) < n) p.log_t++;
  p.lvl_lo = lo;
  p.lvl_hi = hi;
  for (uint32_t k = lo + 1; k <= hi; k++) {
    const Level& lv = levels[k];
    if (!lv.tw_d[0] || !lv.tw_r[1] || !lv.gami[0] || !lv.gam[1] || !lv.gx) return false;
    p.lv[k] = SmallLevel{lv.tw_d[0], lv.tw_r[1], lv.gami[0], lv.gam[1], lv.gx, lv.xnn_s};
  }
  static bool configured = false;
  if (!configured) {
    ECFFT_CUDA(cudaFuncSetAttribute(k_enter_small, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(3 * (1u << LOG_SMALL) * sizeof(Fp))));
    configured = true;
  }
  size_t tiles = (n + ((size_t)1 << p.log_t) - 1) >> p.log_t;
  const bool timed = prof::enabled();
  if (timed) {
    // algorithmic bytes of the depths covered: level passes 64 B/elem each + combines 128 B/elem each
    double bytes = 0;
    for (uint32_t k = lo + 1; k <= hi; k++) bytes += (double)n * (64.0 * 2 * (k - 1) + 128.0);
    prof::record_begin(prof::EXTEND_TILE, bytes, st);
  }
  k_enter_small<<<(unsigned)tiles, 256, 3 * (((size_t)sizeof(Fp)) << p.log_t), st>>>(p);
  if (timed) prof::record_end(st);
  prof::count_launch();
  ECFFT_CUDA(cudaGetLastError());
  return true;
}

}  // namespace k
}  // namespace ecfft
