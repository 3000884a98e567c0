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
  const Fp* dmat;   // MODE 0: decompose matrices | 1: tw_d[source] ({-s1, -s0} per butterfly) | 2: tw_d[source] (1/g)
  const Fp* rmat;   // MODE 0: recombine matrices | 1: tw_r[target] ({s0, s1} per butterfly)   | 2: tw_r[target] (g)
  const Fp* pre;    // MODE 1/2: per-position scale applied on load (1/Gamma^source) or null
  const Fp* post;   // MODE 1/2: per-position scale applied on store (Gamma^target) or null
  unsigned long long nvec;
  unsigned long long total;  // nvec * h, guards the ragged last tile of the packed mode
  uint32_t log_h, j_lo, j_hi, log_c;
  uint32_t log_t;            // tile holds 2^log_t elements
  uint32_t packed;           // 1: h <= tile, a tile is 2^(log_t-log_h) whole consecutive vectors
  uint32_t mode;             // butterfly form: 0 matrix, 1 normalised, 2 symmetric
  uint32_t do_d, do_r, skip_d, skip_r;
};

// Shared-memory tile of Fp.  SOA: the low and the high 16 bytes of the elements live in two separate
// uint4 arrays, so a warp's 16-byte accesses to consecutive elements are bank-conflict free (the
// 32-byte array-of-structures layout is a 2-way conflict on every access).
template <bool SOA>
struct TileMem {
  uint4* s;
  uint32_t T;
  __device__ __forceinline__ Fp ld(uint32_t e) const {
    uint4 a, b;
    if (SOA) { a = s[e]; b = s[T + e]; } else { a = s[2 * e]; b = s[2 * e + 1]; }
    Fp r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
  }
  __device__ __forceinline__ void st(uint32_t e, const Fp& x) const {
    uint4 a = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]), b = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
    if (SOA) { s[e] = a; s[T + e] = b; } else { s[2 * e] = a; s[2 * e + 1] = b; }
  }
};

// One butterfly of the decompose (D = true) or recombine phase on tile elements e_lo, e_hi.
//   MODE 0 — the reference's 2x2 mat-vec (src/utils.rs:338-347): 4 products, 2 lazy reductions; tw -> 4 Fp.
//   MODE 1 — normalised: recombine [[1, s0], [1, s1]] (tw = {s0, s1}); decompose in sum form: with
//            x^_p = -c x_p, x^_q = c x_q (c = 1/(s1-s0), folded into the 1/Gamma pre-scale table)
//            y_q = x^_p + x^_q, y_p = -(s1 x^_p + s0 x^_q) (tw = {-s1, -s0}).  2 products per pair.
//   MODE 2 — symmetric: with g = (s0 - b)/(s0 + b) the pair's two nodes give +g and -g,
//            recombine y_p = x_p + g x_q, y_q = x_p - g x_q (tw = g); decompose x_p = y_p + y_q,
//            x_q = (y_p - y_q)/g (tw = 1/g; the halvings are folded into the pre-scale).  1 product per pair.
template <int MODE, bool D, class TM>
__device__ __forceinline__ void butterfly(const TM& s, uint32_t e_lo, uint32_t e_hi, const Fp* tw) {
  Fp xp = s.ld(e_lo), xq = s.ld(e_hi);
  if (MODE == 0) {
    Fp m0 = fp_load_ro(tw), m1 = fp_load_ro(tw + 1);
    s.st(e_lo, fp_dot2_lazy(m0, xp, m1, xq));
    Fp m2 = fp_load_ro(tw + 2), m3 = fp_load_ro(tw + 3);
    s.st(e_hi, fp_dot2_lazy(m2, xp, m3, xq));
  } else if (MODE == 1) {
    Fp t0 = fp_load_ro(tw), t1 = fp_load_ro(tw + 1);
    if (D) {
      s.st(e_hi, fp_add_lazy(xp, xq));
      s.st(e_lo, fp_dot2_lazy(t0, xp, t1, xq));
    } else {
      s.st(e_lo, fp_muladd_lazy(xp, t0, xq));
      s.st(e_hi, fp_muladd_lazy(xp, t1, xq));
    }
  } else {
    Fp g = fp_load_ro(tw);
    if (D) {
      s.st(e_lo, fp_add_lazy(xp, xq));
      s.st(e_hi, fp_mul_lazy(g, fp_sub_lazy2(xp, xq)));
    } else {
      Fp t = fp_mul_lazy(g, xq);
      s.st(e_lo, fp_add_lazy(xp, t));
      s.st(e_hi, fp_sub_lazy2(xp, t));
    }
  }
}
// Fp entries per butterfly in the twiddle/matrix table, and the table's layer offset for half-stride 2^j
template <int MODE> __device__ __forceinline__ constexpr uint32_t tw_stride() { return MODE == 0 ? 8u : MODE == 1 ? 2u : 1u; }
template <int MODE>
__device__ __forceinline__ const Fp* tw_layer(const Fp* table, uint32_t j, uint32_t skip) {
  return MODE == 0 ? table + 4 * ((2ull << j) + skip) : table + tw_stride<MODE>() * (1ull << j);
}

template <int MODE, int NT, int MINB, bool SOA>
__global__ void __launch_bounds__(NT, MINB) k_extend_tile(TileParams p) {
  extern __shared__ uint4 smem_raw[];
  const uint32_t T = 1u << p.log_t;
  const TileMem<SOA> s{smem_raw, T};
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
    if (MODE != 0 && p.pre) x = fp_mul_lazy(x, fp_load_ro(p.pre + (g & hmask)));
    s.st(e, x);
  }
  __syncthreads();

  if (p.do_d) {
    for (int j = (int)p.j_hi - 1; j >= (int)p.j_lo; j--) {
      const uint32_t sh = (uint32_t)j - p.j_lo + p.log_c;  // bit of the tile index that this level pairs
      const unsigned long long jmask = (1ull << j) - 1;
      const Fp* layer = tw_layer<MODE>(p.dmat, (uint32_t)j, p.skip_d);
#pragma unroll 1
      for (uint32_t b = threadIdx.x; b < T / 2; b += NT) {
        uint32_t e_lo = ((b >> sh) << (sh + 1)) | (b & ((1u << sh) - 1));
        uint32_t r = e_lo >> p.log_c, c = e_lo & (C - 1);
        unsigned long long i = (pos0 + ((unsigned long long)r << p.j_lo) + c) & jmask;
        butterfly<MODE, true>(s, e_lo, e_lo + (1u << sh), layer + tw_stride<MODE>() * i);
      }
      __syncthreads();
    }
  }
  if (p.do_r) {
    for (uint32_t j = p.j_lo; j < p.j_hi; j++) {
      const uint32_t sh = j - p.j_lo + p.log_c;
      const unsigned long long jmask = (1ull << j) - 1;
      const Fp* layer = tw_layer<MODE>(p.rmat, j, p.skip_r);
#pragma unroll 1
      for (uint32_t b = threadIdx.x; b < T / 2; b += NT) {
        uint32_t e_lo = ((b >> sh) << (sh + 1)) | (b & ((1u << sh) - 1));
        uint32_t r = e_lo >> p.log_c, c = e_lo & (C - 1);
        unsigned long long i = (pos0 + ((unsigned long long)r << p.j_lo) + c) & jmask;
        butterfly<MODE, false>(s, e_lo, e_lo + (1u << sh), layer + tw_stride<MODE>() * i);
      }
      __syncthreads();
    }
  }
  for (uint32_t e = threadIdx.x; e < T; e += NT) {
    uint32_t r = e >> p.log_c, c = e & (C - 1);
    unsigned long long g = gbase + ((unsigned long long)r << p.j_lo) + c;
    if (g < p.total) {
      Fp x = s.ld(e);
      if (MODE != 0 && p.post) x = fp_mul_lazy(x, fp_load_ro(p.post + (g & hmask)));
      fp_store(p.out + g, fp_canon(x));
    }
  }
}

// Launch-shape variants of the radix-2 kernel (ECFFT_B200_TILE_VARIANT, measured in profiles/):
//   1: 256 threads, 2048-element tile, __launch_bounds__(256, 2)
//   7: 128 threads, 1024-element tile, __launch_bounds__(128, 6) — default
static int tile_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ECFFT_B200_TILE_VARIANT");
    v = e ? atoi(e) : 7;
    if (v != 1 && v != 7) v = 7;
  }
  return v;
}
static uint32_t log_tile() { return tile_variant() == 1 ? 11 : 10; }

template <int MODE, int NT, int MINB, bool SOA>
static void launch_variant(const TileParams& p, size_t tiles, cudaStream_t st) {
  static PerDeviceOnce configured;
  configured.run([] { ECFFT_CUDA(cudaFuncSetAttribute(k_extend_tile<MODE, NT, MINB, SOA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((1u << 11) * sizeof(Fp)))); });
  k_extend_tile<MODE, NT, MINB, SOA><<<(unsigned)tiles, NT, ((size_t)sizeof(Fp)) << p.log_t, st>>>(p);
}
template <int MODE>
static void launch_mode(const TileParams& p, size_t tiles, cudaStream_t st) {
  if (tile_variant() == 1)
    launch_variant<MODE, 256, 2, true>(p, tiles, st);
  else
    launch_variant<MODE, 128, 6, true>(p, tiles, st);
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
  if (p.mode == 0)
    launch_variant<0, 256, 2, false>(p, tiles, st);
  else if (p.mode == 1)
    launch_mode<1>(p, tiles, st);
  else
    launch_mode<2>(p, tiles, st);
  if (timed) prof::record_end(st);
  prof::count_launch();
  ECFFT_CUDA(cudaGetLastError());
}

// ECFFT_B200_BUTTERFLY = matrix | normalised | symmetric (default; falls back to normalised for trees
// whose rational maps are not of the form (x^2 + c1 x + c0)/x with c0 a square)
int butterfly_mode() {
  static int mode = -1;
  if (mode < 0) {
    const char* e = getenv("ECFFT_B200_BUTTERFLY");
    std::string v = e ? e : "";
    mode = v == "matrix" ? 0 : v == "normalised" ? 1 : 2;
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
  const bool norm = lv.has_norm() && butterfly_mode() != 0;
  if (unscaled_out && !norm) throw Error(ERR_INVALID_ARG, "extend: unscaled output needs the normalised tables");
  const Fp* pre = norm ? lv.gami[source] : nullptr;
  const Fp* post = (norm && !unscaled_out) ? lv.gam[target] : nullptr;
  // symmetric butterflies: the radix-4 register-stage kernel (sym_kernel.cu); the radix-2 tile kernel below
  // remains for the other butterfly forms, for inputs of fewer than 4 elements and ECFFT_B200_SYM_RADIX2
  static const bool radix2 = getenv("ECFFT_B200_SYM_RADIX2") != nullptr;
  if (norm && lv.sym && !radix2 && extend_sym(lv.tw_d[source], lv.tw_r[target], lv.ctr[target], in, out, log_h, nvec, pre, post, nullptr, st)) return;
  p.mode = norm ? (lv.sym ? 2 : 1) : 0;
  p.dmat = norm ? lv.tw_d[source] : lv.dmat;
  p.rmat = norm ? lv.tw_r[target] : lv.rmat;
  p.skip_d = target == S0 ? 1 : 0;  // fftree.rs:87-90
  p.skip_r = target == S1 ? 1 : 0;  // fftree.rs:108-111
  run_passes(p, in, out, log_h, nvec, pre, post, st);
}

// Multi-GPU building block: the rank-local levels (half-strides < 2^log_len) of the normalised
// EXTEND -> S1 of a longer vector whose contiguous chunk of 2^log_len elements this rank holds.
// Twiddles depend only on the position modulo the half-stride, so the chunk behaves like a vector of
// its own length with the long vector's tables; the diagonal scalings are applied by the caller.
void extend_sub(const Level& lv, const Fp* in, Fp* out, uint32_t log_len, cudaStream_t st, const Fp* pre, Moiety source, Moiety target) {
  if (!lv.has_norm()) throw Error(ERR_MISSING_TABLES, "extend_sub: normalised tables missing");
  if (log_len == 0) {
    if (pre) mul_bcast(out, in, pre, 1, 1, st);
    else if (in != out) ECFFT_CUDA(cudaMemcpyAsync(out, in, sizeof(Fp), cudaMemcpyDeviceToDevice, st));
    return;
  }
  if (lv.sym && extend_sym(lv.tw_d[source], lv.tw_r[target], lv.ctr[target], in, out, log_len, 1, pre, nullptr, nullptr, st)) return;
  TileParams p;
  p.mode = lv.sym ? 2 : 1;
  p.dmat = lv.tw_d[source];
  p.rmat = lv.tw_r[target];
  p.skip_d = target == S0 ? 1 : 0;
  p.skip_r = target == S1 ? 1 : 0;
  run_passes(p, in, out, log_len, 1, pre, nullptr, st);
}

}  // namespace k
}  // namespace ecfft
