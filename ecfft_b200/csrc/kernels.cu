// sm_100a kernels of the ECFFT engine.
//
// Hot kernels (named, profiled):
//   k_extend_tile    — several consecutive EXTEND butterfly levels (reference
//                      src/fftree.rs:81-97 decompose, :103-118 recombine; butterfly
//                      src/utils.rs:338-347) on a tile held in shared memory.
//   k_enter_combine  — ENTER's u + v * x^(n/2) interleave (src/fftree.rs:155-159).
// Everything else is O(n) pointwise glue launched through a generic grid-stride kernel.
#include "engine.h"
#include "ec.cuh"

#include <atomic>
#include <cstdlib>
#include <mutex>
#include <string>

namespace ecfft {

namespace prof {
struct Rec { Kernel k; double bytes; cudaEvent_t e0, e1; };
static std::atomic<unsigned long long> g_launches{0};
static std::atomic<bool> g_enabled{false};
static std::mutex g_mu;
static std::vector<Rec> g_recs;
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }
unsigned long long launches() { return g_launches.load(); }
void enable(bool on) { g_enabled.store(on); }
bool enabled() { return g_enabled.load(); }
void record_begin(Kernel k, double alg_bytes, cudaStream_t st) {
  Rec r;
  r.k = k;
  r.bytes = alg_bytes;
  ECFFT_CUDA(cudaEventCreate(&r.e0));
  ECFFT_CUDA(cudaEventCreate(&r.e1));
  ECFFT_CUDA(cudaEventRecord(r.e0, st));
  std::lock_guard<std::mutex> lock(g_mu);
  g_recs.push_back(r);
}
void record_end(cudaStream_t st) {
  std::lock_guard<std::mutex> lock(g_mu);
  ECFFT_CUDA(cudaEventRecord(g_recs.back().e1, st));
}
void read(Kernel k, double* ms, double* alg_bytes, unsigned long long* n) {
  std::lock_guard<std::mutex> lock(g_mu);
  *ms = 0;
  *alg_bytes = 0;
  *n = 0;
  std::vector<Rec> keep;
  for (Rec& r : g_recs) {
    if (r.k != k) {
      keep.push_back(r);
      continue;
    }
    float t = 0;
    ECFFT_CUDA(cudaEventSynchronize(r.e1));
    ECFFT_CUDA(cudaEventElapsedTime(&t, r.e0, r.e1));
    *ms += t;
    *alg_bytes += r.bytes;
    *n += 1;
    cudaEventDestroy(r.e0);
    cudaEventDestroy(r.e1);
  }
  g_recs.swap(keep);
}
}  // namespace prof

namespace k {

static constexpr int NT = 256;              // threads per CTA for the tile kernel
static constexpr uint32_t LOG_TILE = 11;    // 2048 elements = 64 KiB of shared memory per CTA

static inline unsigned grid_for(size_t n, int threads) {
  size_t g = (n + threads - 1) / threads;
  size_t cap = 148u * 64u;  // grid-stride beyond this
  if (g > cap) g = cap;
  if (g == 0) g = 1;
  return (unsigned)g;
}

template <class F>
__global__ void __launch_bounds__(256) k_map(size_t n, F f) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) f(i);
}
template <class F>
static void map(size_t n, cudaStream_t st, F f) {
  if (n == 0) return;
  k_map<<<grid_for(n, 256), 256, 0, st>>>(n, f);
  prof::count_launch();
  ECFFT_CUDA(cudaGetLastError());
}

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
__device__ __forceinline__ void butterfly_norm_r(Fp* s, uint32_t e_lo, uint32_t e_hi, const Fp* tw) {
  Fp s0 = fp_load_ro(tw), s1 = fp_load_ro(tw + 1);
  Fp xp = s[e_lo], xq = s[e_hi];
  s[e_lo] = fp_muladd_lazy(xp, s0, xq);
  s[e_hi] = fp_muladd_lazy(xp, s1, xq);
}
// MODE 1 decompose: inverse of [[1, s0], [1, s1]]: y_q = (x_q - x_p)/(s1 - s0), y_p = x_p - s0*y_q
__device__ __forceinline__ void butterfly_norm_d(Fp* s, uint32_t e_lo, uint32_t e_hi, const Fp* tw) {
  Fp c = fp_load_ro(tw), ns0 = fp_load_ro(tw + 1);
  Fp xp = fp_canon(s[e_lo]), xq = s[e_hi];
  Fp yq = fp_mul_lazy(c, fp_sub_lazy(xq, xp));
  s[e_hi] = yq;
  s[e_lo] = fp_muladd_lazy(xp, ns0, yq);
}

template <int MODE>
__global__ void __launch_bounds__(NT, 2) k_extend_tile(TileParams p) {
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
      for (uint32_t b = threadIdx.x; b < T / 2; b += NT) {
        uint32_t e_lo = ((b >> sh) << (sh + 1)) | (b & ((1u << sh) - 1));
        uint32_t r = e_lo >> p.log_c, c = e_lo & (C - 1);
        unsigned long long i = (pos0 + ((unsigned long long)r << p.j_lo) + c) & jmask;
        if (MODE == 0)
          butterfly_matrix(s, e_lo, e_lo + (1u << sh), layer + 8 * i);
        else
          butterfly_norm_d(s, e_lo, e_lo + (1u << sh), layer + 2 * i);
      }
      __syncthreads();
    }
  }
  if (p.do_r) {
    for (uint32_t j = p.j_lo; j < p.j_hi; j++) {
      const uint32_t sh = j - p.j_lo + p.log_c;
      const unsigned long long jmask = (1ull << j) - 1;
      const Fp* layer = MODE == 0 ? p.rmat + 4 * ((2ull << j) + p.skip_r) : p.rmat + 2 * (1ull << j);
      for (uint32_t b = threadIdx.x; b < T / 2; b += NT) {
        uint32_t e_lo = ((b >> sh) << (sh + 1)) | (b & ((1u << sh) - 1));
        uint32_t r = e_lo >> p.log_c, c = e_lo & (C - 1);
        unsigned long long i = (pos0 + ((unsigned long long)r << p.j_lo) + c) & jmask;
        if (MODE == 0)
          butterfly_matrix(s, e_lo, e_lo + (1u << sh), layer + 8 * i);
        else
          butterfly_norm_r(s, e_lo, e_lo + (1u << sh), layer + 2 * i);
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

static void launch_tile(const TileParams& p, cudaStream_t st) {
  static bool configured = false;
  if (!configured) {
    ECFFT_CUDA(cudaFuncSetAttribute(k_extend_tile<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((1u << LOG_TILE) * sizeof(Fp))));
    ECFFT_CUDA(cudaFuncSetAttribute(k_extend_tile<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((1u << LOG_TILE) * sizeof(Fp))));
    configured = true;
  }
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
  if (p.norm)
    k_extend_tile<1><<<(unsigned)tiles, NT, ((size_t)sizeof(Fp)) << p.log_t, st>>>(p);
  else
    k_extend_tile<0><<<(unsigned)tiles, NT, ((size_t)sizeof(Fp)) << p.log_t, st>>>(p);
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
  const Fp* pre = norm ? lv.gami[source] : nullptr;
  const Fp* post = (norm && !unscaled_out) ? lv.gam[target] : nullptr;
  p.pre = nullptr;
  p.post = nullptr;
  p.nvec = nvec;
  p.total = nvec << log_h;
  p.log_h = log_h;
  p.log_t = LOG_TILE;
  p.packed = 0;
  p.skip_d = target == S0 ? 1 : 0;  // fftree.rs:87-90
  p.skip_r = target == S1 ? 1 : 0;  // fftree.rs:108-111
  if (log_h <= LOG_TILE) {
    // whole vectors fit a tile: pack 2^(log_t-log_h) consecutive vectors per CTA
    p.in = in; p.out = out;
    p.j_lo = 0; p.j_hi = log_h; p.log_c = 0; p.do_d = 1; p.do_r = 1;
    p.packed = 1;
    p.log_t = log_h;
    while (p.log_t < LOG_TILE && ((size_t)1 << p.log_t) < p.total) p.log_t++;
    p.pre = pre;
    p.post = post;
    launch_tile(p, st);
    return;
  }
  // outer decompose passes (strided tiles), inner fused pass, outer recombine passes
  const uint32_t outer = log_h - LOG_TILE;
  const uint32_t kmax = 6;                        // 64 rows x 32 columns (1 KiB contiguous per row)
  const uint32_t npass = (outer + kmax - 1) / kmax;
  std::vector<uint32_t> bounds;                   // j boundaries from log_h down to LOG_TILE
  bounds.push_back(log_h);
  for (uint32_t i = 1; i <= npass; i++) bounds.push_back(log_h - (outer * i) / npass);
  const Fp* src = in;
  for (uint32_t i = 0; i < npass; i++) {
    p.in = src; p.out = out;
    p.j_hi = bounds[i]; p.j_lo = bounds[i + 1]; p.log_c = LOG_TILE - (p.j_hi - p.j_lo);
    p.do_d = 1; p.do_r = 0;
    p.pre = i == 0 ? pre : nullptr;  // 1/Gamma^source on the very first load
    launch_tile(p, st);
    src = out;
  }
  p.pre = nullptr;
  p.in = out; p.out = out; p.j_lo = 0; p.j_hi = LOG_TILE; p.log_c = 0; p.do_d = 1; p.do_r = 1;
  launch_tile(p, st);
  for (uint32_t i = npass; i-- > 0;) {
    p.in = out; p.out = out;
    p.j_hi = bounds[i]; p.j_lo = bounds[i + 1]; p.log_c = LOG_TILE - (p.j_hi - p.j_lo);
    p.do_d = 0; p.do_r = 1;
    p.post = i == 0 ? post : nullptr;  // Gamma^target on the very last store
    launch_tile(p, st);
  }
}

// ------------------------------------------------------------------------------------------
// ENTER combine: res[2i] = u0[i] + v0[i]*xnn[2i], res[2i+1] = u1[i] + v1[i]*xnn[2i+1]
// A holds [u0 | v0] per block of 2h, W holds [u1 | v1].  SCALED: W lacks the Gamma^1 scaling of the
// normalised EXTEND, so res[2i+1] = gam[i]*u1^[i] + (gam[i]*xnn[2i+1])*v1^[i] with both constants
// precomputed (one lazy reduction for the two products).
// ------------------------------------------------------------------------------------------
template <bool SCALED>
__global__ void __launch_bounds__(256) k_enter_combine(const Fp* __restrict__ A, const Fp* __restrict__ W,
                                                       const Fp* __restrict__ xnn, const Fp* __restrict__ gam,
                                                       const Fp* __restrict__ gx, Fp* __restrict__ out,
                                                       uint32_t log_h, unsigned long long npairs) {
  for (unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; idx < npairs;
       idx += (unsigned long long)gridDim.x * blockDim.x) {
    unsigned long long blk = idx >> log_h, i = idx & ((1ull << log_h) - 1);
    unsigned long long off = blk << (log_h + 1);
    unsigned long long h = 1ull << log_h;
    Fp u0 = fp_load(A + off + i), v0 = fp_load(A + off + h + i);
    Fp x0 = fp_load_ro(xnn + 2 * i);
    Fp r0 = fp_canon(fp_muladd_lazy(u0, v0, x0));
    fp_store(out + off + 2 * i, r0);
    Fp u1 = fp_load(W + off + i), v1 = fp_load(W + off + h + i);
    Fp r1;
    if (SCALED) {
      r1 = fp_canon(fp_dot2_lazy(fp_load_ro(gam + i), u1, fp_load_ro(gx + i), v1));
    } else {
      Fp x1 = fp_load_ro(xnn + 2 * i + 1);
      r1 = fp_canon(fp_muladd_lazy(u1, v1, x1));
    }
    fp_store(out + off + 2 * i + 1, r1);
  }
}
void enter_combine(const Level& lv, const Fp* A, const Fp* W, Fp* out, uint32_t log_h, size_t n, bool W_unscaled, cudaStream_t st) {
  size_t npairs = n / 2;
  unsigned grid = (unsigned)((npairs + 255) / 256);
  if (grid > 148u * 32u) grid = 148u * 32u;
  const bool timed = prof::enabled();
  if (timed) prof::record_begin(prof::ENTER_COMBINE, 128.0 * (double)n, st);  // u0,v0,u1,v1 / xnn / out
  if (W_unscaled)
    k_enter_combine<true><<<grid, 256, 0, st>>>(A, W, lv.xnn_s, lv.gam[1], lv.gx, out, log_h, npairs);
  else
    k_enter_combine<false><<<grid, 256, 0, st>>>(A, W, lv.xnn_s, nullptr, nullptr, out, log_h, npairs);
  if (timed) prof::record_end(st);
  prof::count_launch();
  ECFFT_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------
// pointwise glue
// ------------------------------------------------------------------------------------------
void mul_const(Fp* out, const Fp* in, Fp c, size_t n, cudaStream_t st) {
  map(n, st, [=] __device__(size_t i) { fp_store(out + i, fp_mul(fp_load(in + i), c)); });
}
void mul_bcast(Fp* out, const Fp* in, const Fp* c, size_t len, size_t nvec, cudaStream_t st) {
  map(len * nvec, st, [=] __device__(size_t i) { fp_store(out + i, fp_mul(fp_load(in + i), fp_load_ro(c + i % len))); });
}
void add_bcast_scaled(Fp* out, const Fp* in, const Fp* z, Fp scale, size_t len, size_t nvec, cudaStream_t st) {
  map(len * nvec, st, [=] __device__(size_t i) {
    fp_store(out + i, fp_canon(fp_muladd_lazy(fp_load(in + i), fp_load_ro(z + i % len), scale)));
  });
}
void deinterleave(Fp* even, Fp* odd, const Fp* in, size_t pairs, cudaStream_t st) {
  map(pairs, st, [=] __device__(size_t i) {
    fp_store(even + i, fp_load(in + 2 * i));
    fp_store(odd + i, fp_load(in + 2 * i + 1));
  });
}
void interleave(Fp* out, const Fp* even, const Fp* odd, size_t pairs, cudaStream_t st) {
  map(pairs, st, [=] __device__(size_t i) {
    fp_store(out + 2 * i, fp_load(even + i));
    fp_store(out + 2 * i + 1, fp_load(odd + i));
  });
}
void copy_strided(Fp* out, const Fp* in, size_t count, size_t in_stride, cudaStream_t st) {
  map(count, st, [=] __device__(size_t i) { fp_store(out + i, fp_load(in + i * in_stride)); });
}
void fill(Fp* out, Fp c, size_t n, cudaStream_t st) {
  map(n, st, [=] __device__(size_t i) { fp_store(out + i, c); });
}
// t0[v][i] = evals[v][2i] * a0inv[i]
void redc_pre(Fp* t0, const Fp* evals, const Fp* a0inv, size_t h, size_t nvec, cudaStream_t st) {
  map(h * nvec, st, [=] __device__(size_t idx) {
    size_t v = idx / h, i = idx % h;
    fp_store(t0 + idx, fp_mul(fp_load(evals + v * 2 * h + 2 * i), fp_load_ro(a0inv + i)));
  });
}
// h1[v][i] = (evals[v][2i+1] - g1[v][i] * a[2i+1]) * zinv[i]
void redc_mid(Fp* h1, const Fp* evals, const Fp* g1, const Fp* a, const Fp* zinv, size_t h, size_t nvec, cudaStream_t st) {
  map(h * nvec, st, [=] __device__(size_t idx) {
    size_t v = idx / h, i = idx % h;
    Fp ga = fp_mul(fp_load(g1 + idx), fp_load_ro(a + 2 * i + 1));
    Fp d = fp_sub(fp_load(evals + v * 2 * h + 2 * i + 1), ga);
    fp_store(h1 + idx, fp_mul(d, fp_load_ro(zinv + i)));
  });
}
void exit_split(Fp* next, const Fp* evals, const Fp* M, const Fp* xnn_inv, size_t h, size_t nvec, cudaStream_t st) {
  map(h * nvec, st, [=] __device__(size_t idx) {
    size_t v = idx / h, i = idx % h;
    Fp u0 = fp_load(M + v * 2 * h + 2 * i);
    Fp e0 = fp_load(evals + v * 2 * h + 2 * i);
    Fp v0 = fp_mul(fp_sub(e0, u0), fp_load_ro(xnn_inv + 2 * i));
    fp_store(next + v * 2 * h + i, u0);
    fp_store(next + v * 2 * h + h + i, v0);
  });
}
void vanish_base(Fp* out, const Fp* dom, Fp l0, Fp l1, size_t n, cudaStream_t st) {
  map(n, st, [=] __device__(size_t i) {
    Fp a = fp_load(dom + i);
    fp_store(out + 2 * i, fp_sub(a, l0));
    fp_store(out + 2 * i + 1, fp_sub(a, l1));
  });
}
void mul_pairs(Fp* q0, const Fp* Q, size_t len, size_t npairs, int fix_mont, cudaStream_t st) {
  map(len * npairs, st, [=] __device__(size_t idx) {
    size_t w = idx / len, i = idx % len;
    Fp r = fp_mul(fp_load(Q + 2 * w * len + i), fp_load(Q + (2 * w + 1) * len + i));
    if (fix_mont) r = fp_mul(r, fp_const_RINV());
    fp_store(q0 + idx, r);
  });
}
// out[w][2i] = q0[w][i]; out[w][2i+1] = e[w][i] + z[i]*zscale
void vanish_merge(Fp* out, const Fp* q0, const Fp* e, const Fp* z, Fp zscale, size_t len, size_t nvec, cudaStream_t st) {
  map(len * nvec, st, [=] __device__(size_t idx) {
    size_t w = idx / len, i = idx % len;
    fp_store(out + w * 2 * len + 2 * i, fp_load(q0 + idx));
    fp_store(out + w * 2 * len + 2 * i + 1, fp_canon(fp_muladd_lazy(fp_load(e + idx), fp_load_ro(z + i), zscale)));
  });
}
void count_neq(unsigned long long* counter, const Fp* a, const Fp* b, size_t n, cudaStream_t st) {
  map(n, st, [=] __device__(size_t i) {
    if (!fp_eq(fp_load(a + i), fp_load(b + i))) atomicAdd(counter, 1ull);
  });
}
void sub_mul_bcast(Fp* out, const Fp* a, const Fp* b, const Fp* c, size_t len, size_t nvec, cudaStream_t st) {
  map(len * nvec, st, [=] __device__(size_t i) {
    fp_store(out + i, fp_mul(fp_sub(fp_load(a + i), fp_load(b + i)), fp_load_ro(c + i % len)));
  });
}
void pow_u64(Fp* out, const Fp* in, uint64_t e, size_t n, cudaStream_t st) {
  map(n, st, [=] __device__(size_t i) { fp_store(out + i, fp_pow_u64(fp_load(in + i), e)); });
}
void count_noncanonical(unsigned long long* counter, const Fp* v, size_t n, cudaStream_t st) {
  map(n, st, [=] __device__(size_t i) {
    Fp x = fp_load(v + i);
    if (!fp_eq(x, fp_canon(x))) atomicAdd(counter, 1ull);
  });
}
// q[i] = ((z_i - xnn[i])^2 - sub[i]) * mul[i], z_i = z_half[i/2] when (i&1)==z_parity else 0;
// sub/mul may be null (then q = (z_i - xnn[i])^2).  fftree.rs:430-438, 449-451
void sqr_sub_mul(Fp* out, const Fp* z_half, int z_parity, const Fp* xnn, const Fp* sub, const Fp* mul, size_t n, cudaStream_t st) {
  map(n, st, [=] __device__(size_t i) {
    Fp z = ((int)(i & 1) == z_parity) ? fp_load(z_half + i / 2) : fp_zero();
    Fp d = fp_sub(z, fp_load(xnn + i));
    Fp q = fp_mul(d, d);
    if (sub) q = fp_sub(q, fp_load(sub + i));
    if (mul) q = fp_mul(q, fp_load(mul + i));
    fp_store(out + i, q);
  });
}
void muladd(Fp* out, const Fp* a, const Fp* b, const Fp* c, size_t n, cudaStream_t st) {
  map(n, st, [=] __device__(size_t i) {
    fp_store(out + i, fp_canon(fp_muladd_lazy(fp_load(a + i), fp_load(b + i), fp_load(c + i))));
  });
}

// ------------------------------------------------------------------------------------------
// batch inversion (ark_ff::batch_inversion semantics: zeros untouched).  Each thread owns
// KINV elements (Montgomery's trick), a warp shares one Fermat inversion through prefix and
// suffix product scans over shuffles: ~9 multiplications per element.
// ------------------------------------------------------------------------------------------
static constexpr int KINV = 4;
__device__ __forceinline__ Fp fp_shfl(const Fp& x, int src) {
  Fp r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = __shfl_sync(0xffffffffu, x.v[i], src);
  return r;
}
__global__ void __launch_bounds__(128) k_batch_inverse(Fp* v, size_t n) {
  const int lane = threadIdx.x & 31;
  const size_t warp = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const size_t nwarps = ((size_t)gridDim.x * blockDim.x) >> 5;
  for (size_t base = warp * (32 * KINV); base < n; base += nwarps * (32 * KINV)) {
    Fp x[KINV], pre[KINV];
    bool zero[KINV];
#pragma unroll
    for (int j = 0; j < KINV; j++) {
      size_t idx = base + lane + 32 * j;
      x[j] = idx < n ? fp_load(v + idx) : fp_one();
      zero[j] = fp_is_zero(x[j]);
      if (zero[j]) x[j] = fp_one();
      pre[j] = j == 0 ? x[0] : fp_mul(pre[j - 1], x[j]);
    }
    Fp total = pre[KINV - 1];
    Fp P = total, S = total;  // inclusive prefix / suffix products of the lane totals
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      Fp y = fp_shfl(P, lane - d < 0 ? 0 : lane - d);
      if (lane >= d) P = fp_mul(P, y);
      Fp z = fp_shfl(S, lane + d > 31 ? 31 : lane + d);
      if (lane + d <= 31) S = fp_mul(S, z);
    }
    Fp winv = fp_inv(fp_shfl(P, 31));
    Fp pe = fp_shfl(P, lane == 0 ? 0 : lane - 1);
    Fp se = fp_shfl(S, lane == 31 ? 31 : lane + 1);
    Fp inv = winv;  // becomes 1/total of this lane
    if (lane > 0) inv = fp_mul(inv, pe);
    if (lane < 31) inv = fp_mul(inv, se);
#pragma unroll
    for (int j = KINV - 1; j >= 0; j--) {
      Fp r = j == 0 ? inv : fp_mul(inv, pre[j - 1]);
      if (j > 0) inv = fp_mul(inv, x[j]);
      size_t idx = base + lane + 32 * j;
      if (idx < n && !zero[j]) fp_store(v + idx, r);
    }
  }
}
void batch_inverse(Fp* v, size_t n, cudaStream_t st) {
  if (n == 0) return;
  size_t warps = (n + 32 * KINV - 1) / (32 * KINV);
  size_t blocks = (warps + 3) / 4;
  if (blocks > 148 * 16) blocks = 148 * 16;
  k_batch_inverse<<<(unsigned)blocks, 128, 0, st>>>(v, n);
  prof::count_launch();
  ECFFT_CUDA(cudaGetLastError());
}

// ------------------------------------------------------------------------------------------
// tree construction kernels (plain-form values)
// ------------------------------------------------------------------------------------------
static constexpr int LEAF_CHUNK = 16;
// leaves[i] = x(offset + i*G), reference src/lib.rs:72-78; gtab[j] = 2^j * G (x,y pairs)
__global__ void __launch_bounds__(128) k_build_leaves(Fp* leaves, size_t n, Fp a, Fp a4, Fp offx, Fp offy, const Fp* gtab, uint32_t log_n) {
  size_t t = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t base = t * LEAF_CHUNK;
  if (base >= n) return;
  Pt acc = pt_infinity();
  for (uint32_t j = 0; j < log_n; j++)
    if ((base >> j) & 1) {
      Pt g;
      g.inf = false;
      g.x = fp_load(gtab + 2 * j);
      g.y = fp_load(gtab + 2 * j + 1);
      acc = pt_add(acc, g, a, a4);
    }
  Pt off, g0;
  off.inf = false; off.x = offx; off.y = offy;
  g0.inf = log_n == 0;
  if (!g0.inf) { g0.x = fp_load(gtab); g0.y = fp_load(gtab + 1); }
  Pt p = pt_add(off, acc, a, a4);
  for (int c = 0; c < LEAF_CHUNK && base + c < n; c++) {
    fp_store(leaves + base + c, p.x);
    if (c + 1 < LEAF_CHUNK && base + c + 1 < n) p = pt_add(p, g0, a, a4);
  }
}
void build_leaves(Fp* leaves, size_t n, Fp a, Fp a4, Fp offx, Fp offy, const Fp* gtab_xy, uint32_t log_n, cudaStream_t st) {
  size_t threads = (n + LEAF_CHUNK - 1) / LEAF_CHUNK;
  k_build_leaves<<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(leaves, n, a, a4, offx, offy, gtab_xy, log_n);
  prof::count_launch();
  ECFFT_CUDA(cudaGetLastError());
}

__device__ __forceinline__ Fp poly_eval_dev(const Fp* c, int n, const Fp& x) {
  Fp acc = fp_zero();
  for (int i = n - 1; i >= 0; i--) acc = fp_add(fp_mul(acc, x), fp_load_ro(c + i));
  return acc;
}
// layer[i] = num(prev[i]) / den(prev[i])   (RationalMap::map, reference src/utils.rs:383-385)
void ratmap_layer(Fp* layer, const Fp* prev, size_t count, const Fp* num, int nnum, const Fp* den, int nden, cudaStream_t st) {
  map(count, st, [=] __device__(size_t i) {
    Fp x = fp_load(prev + i);
    Fp nu = poly_eval_dev(num, nnum, x), de = poly_eval_dev(den, nden, x);
    fp_store(layer + i, fp_mul(nu, fp_inv(de)));
  });
}
// Lemma 3.2 matrices of one layer, reference src/fftree.rs:354-362.  flayer has 2d entries at
// stride fstride (the chain level's f layer is a strided view of the top tree's).
void build_matrices(Fp* rl, Fp* dl, const Fp* flayer, size_t fstride, size_t d, const Fp* den, int nden, cudaStream_t st) {
  uint64_t e = d / 2 - 1;
  map(d, st, [=] __device__(size_t i) {
    Fp s0 = fp_load(flayer + i * fstride), s1 = fp_load(flayer + (i + d) * fstride);
    Fp v0 = fp_pow_u64(poly_eval_dev(den, nden, s0), e);
    Fp v1 = fp_pow_u64(poly_eval_dev(den, nden, s1), e);
    Fp r0 = v0, r1 = fp_mul(s0, v0), r2 = v1, r3 = fp_mul(s1, v1);
    Fp det = fp_sub(fp_mul(r0, r3), fp_mul(r1, r2));
    Fp di = fp_inv(det);
    Fp* r = rl + 4 * i;
    Fp* m = dl + 4 * i;
    fp_store(r, r0); fp_store(r + 1, r1); fp_store(r + 2, r2); fp_store(r + 3, r3);
    fp_store(m, fp_mul(r3, di));
    fp_store(m + 1, fp_mul(fp_neg(r1), di));
    fp_store(m + 2, fp_mul(fp_neg(r2), di));
    fp_store(m + 3, fp_mul(r0, di));
  });
}

// Normalised-butterfly tables of one chain level (DESIGN.md "twiddle form").  Entry idx = 2^j + i:
// s0 = f[2B + 2i + mu], s1 = f[2B + 2i + mu + B] with B = 2^(j+1) — the same nodes the reference's
// matrices are built from (src/fftree.rs:356-357) — through the strided view of the top tree's f.
void build_twiddles(Fp* tw_r, Fp* tw_d, const Fp* f_top, size_t fstride, size_t h, int mu, cudaStream_t st) {
  map(h, st, [=] __device__(size_t idx) {
    if (idx == 0) {
      fp_store(tw_r, fp_zero()); fp_store(tw_r + 1, fp_zero());
      fp_store(tw_d, fp_zero()); fp_store(tw_d + 1, fp_zero());
      return;
    }
    uint32_t j = 63 - __clzll((unsigned long long)idx);
    size_t i = idx - ((size_t)1 << j), B = (size_t)2 << j;
    Fp s0 = fp_load(f_top + (2 * B + 2 * i + mu) * fstride);
    Fp s1 = fp_load(f_top + (2 * B + 2 * i + mu + B) * fstride);
    fp_store(tw_r + 2 * idx, s0);
    fp_store(tw_r + 2 * idx + 1, s1);
    fp_store(tw_d + 2 * idx, fp_inv(fp_sub(s1, s0)));
    fp_store(tw_d + 2 * idx + 1, fp_neg(s0));
  });
}
// Gamma^mu_p = prod_j v(node_j(p))^(2^j - 1): exactly the first-column entries of the recombine
// matrices the position passes through (R = [[v0, s0 v0], [v1, s1 v1]], src/fftree.rs:360)
void build_gamma(Fp* gam, const Fp* rmat, size_t h, int mu, cudaStream_t st) {
  map(h, st, [=] __device__(size_t p) {
    Fp acc = fp_one();
    for (uint32_t j = 0; ((size_t)1 << j) < h; j++) {
      size_t i = p & (((size_t)1 << j) - 1), b = (p >> j) & 1;
      acc = fp_mul_lazy(acc, fp_load(rmat + 4 * (((size_t)2 << j) + 2 * i + mu) + 2 * b));
    }
    fp_store(gam + p, fp_canon(acc));
  });
}
void mul_strided(Fp* out, const Fp* a, const Fp* b, size_t b_stride, size_t b_off, size_t n, cudaStream_t st) {
  map(n, st, [=] __device__(size_t i) { fp_store(out + i, fp_mul(fp_load(a + i), fp_load(b + b_off + i * b_stride))); });
}

}  // namespace k
}  // namespace ecfft
