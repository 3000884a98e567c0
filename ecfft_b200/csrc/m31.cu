// FFTree over the Mersenne-31 field (reference `ecfft::m31`, src/lib.rs:190-215): the reference's second field,
// SURVEY.md 8f.4.  The same eight algorithms (src/fftree.rs:72-316) and the same construction (build_ec_fftree
// src/ec.rs:498-554, Velu 2-isogenies src/ec.rs:214-243, FFTree::new / from_tree src/fftree.rs:42-70, 318-463) as the
// secp256k1 engine, for 4-byte elements: a u32 holding the canonical value, which is what the reference's
// `ark_ff_optimized::fp31::Fp` keeps in memory.  The rational maps here are (x^2 - x0 x + t)/(x - x0), not the
// Good-curve shape, so the butterflies are the reference's 2x2 matrices (src/utils.rs:338-347); with 4-byte
// elements the path is HBM-bound on its matrix tables and a level costs four 32x32 products per pair.
//
// Layout: vectors are contiguous u32 arrays; per chain level N the matrices are uint4 (row major m00 m01 m10 m11)
// in the reference's BinaryTree order (entry 2^(j+1) + 2i + skip for butterfly i of the level with half-stride 2^j).
// One CTA keeps a tile of 4096 elements (16 KiB) in shared memory for a group of consecutive levels — contiguous
// for the innermost 12 + 12 levels, rows of >= 128 contiguous elements for the outer ones — exactly the pass
// structure of the secp256k1 kernel, radix 2.
#include <array>
#include <cstring>
#include <memory>

#include "engine.h"
#include "../../include/ecfft_b200.h"

namespace ecfft {
namespace m31 {

typedef uint32_t F;
static constexpr uint32_t P31 = 0x7fffffffu;

__host__ __device__ __forceinline__ F fadd(F a, F b) { uint32_t s = a + b; return s >= P31 ? s - P31 : s; }
__host__ __device__ __forceinline__ F fsub(F a, F b) { return a >= b ? a - b : a + P31 - b; }
__host__ __device__ __forceinline__ F fred(uint64_t t) {   // t < 2^63
  uint64_t s = (t & P31) + (t >> 31);                       // < 2^33
  uint32_t r = (uint32_t)(s & P31) + (uint32_t)(s >> 31);   // < 2^31 + 4
  return r >= P31 ? r - P31 : r;
}
__host__ __device__ __forceinline__ F fmul(F a, F b) { return fred((uint64_t)a * b); }
__host__ __device__ __forceinline__ F fdot2(F a, F b, F c, F d) { return fred((uint64_t)a * b + (uint64_t)c * d); }
__host__ __device__ __forceinline__ F fneg(F a) { return a ? P31 - a : 0; }
__host__ __device__ inline F fpow(F x, uint64_t e) {
  F r = 1;
  while (e) {
    if (e & 1) r = fmul(r, x);
    x = fmul(x, x);
    e >>= 1;
  }
  return r;
}
__host__ __device__ inline F finv(F x) { return fpow(x, P31 - 2); }   // 0 -> 0 (ark_ff::batch_inversion leaves zeros)

// ---------------------------------------------------------------------------------------------------------
// device kernels
// ---------------------------------------------------------------------------------------------------------
template <class Fn>
__global__ void k31_map(size_t n, Fn fn) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // programmatic dependent launch, see k_extend_sym
  asm volatile("griddepcontrol.wait;" ::: "memory");
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) fn(i);
}
// Programmatic dependent launch is decided per CALL, for all of its launches: measured (profiles/r02_ab_m31_pdl.txt) it
// gains 15-20 % at n = 2^16 (launch gaps dominate), is mixed at 2^20 and loses at 2^22 (the next grid's early-resident CTAs
// take slots from the running one); attaching it to some launches of a call and not to others is the worst of all
// (EXIT 2^22: 7.9 ms without, 8.9 ms on every launch, 10.6 ms on the small grids only).
// ECFFT_B200_M31_PDL: 0 = never, 1 = calls of at most 2^20 elements (default), 2 = always.
static int pdl_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ECFFT_B200_M31_PDL");
    v = e ? atoi(e) : 1;
  }
  return v;
}
static thread_local bool t_pdl = false;   // set by the entry point for the call it runs
static void set_call_size(size_t n) { t_pdl = pdl_mode() == 2 || (pdl_mode() == 1 && n <= ((size_t)1 << 20)); }
template <class Fn>
static void map(size_t n, cudaStream_t st, Fn fn) {
  if (!n) return;
  size_t blocks = (n + 255) / 256;
  if (blocks > 148u * 16u) blocks = 148u * 16u;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)blocks);
  cfg.blockDim = dim3(256);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = t_pdl ? 1 : 0;
  ECFFT_CUDA(cudaLaunchKernelEx(&cfg, k31_map<Fn>, n, fn));
  prof::count_launch();
}

static constexpr uint32_t LT = 12;   // log2 tile elements
struct Pass {
  const F* in;
  F* out;
  const uint4* dmat;
  const uint4* rmat;
  const F* tw_d;    // symmetric form: 1/g of the source moiety, entry 2^j + i
  const F* tw_r;    // symmetric form: g of the target moiety
  const F* pre;     // per-position scale applied as the tile is loaded (or null)
  const F* post;    // per-position scale applied as the tile is stored (or null)
  unsigned long long total, nv;
  uint32_t log_h, lvl_lo, lvl_hi, do_d, do_r, dskip, rskip;
  uint32_t packed, log_t, log_c, krows, row_shift;
};

// One pass of EXTEND on a tile (flattening of extend_impl, src/fftree.rs:72-120: decompose levels with
// half-strides 2^(lvl_hi-1) .. 2^lvl_lo, then recombine levels back up), in place in shared memory.
// SYM = false: the reference's 2x2 matrices (src/utils.rs:338-347), four products per pair.
// SYM = true: the one-product butterflies of DESIGN.md 4.1 — every map of the m31 chain is, in the coordinate
// y = x - x0, y + t/y + x0 with t a square (src/ec.rs:231-232), so the two nodes of a pair are y and beta^2/y and
// g = (y - beta)/(y + beta) takes opposite values on them: recombine y_p = x_p + g x_q, y_q = x_p - g x_q, decompose
// x_p = y_p + y_q, x_q = (y_p - y_q)/g, the diagonal scalings collected into one pre- and one post-scale.
template <bool SYM>
__global__ void __launch_bounds__(256) k31_extend(const __grid_constant__ Pass p) {
  extern __shared__ F tile[];
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const uint32_t T = 1u << p.log_t;
  // tile element e -> global element and position within its vector
  unsigned long long gbase;
  uint32_t pos0 = 0;
  if (p.packed) {
    gbase = (unsigned long long)blockIdx.x << p.log_t;
  } else {
    const unsigned long long w = blockIdx.x % p.nv, tl = blockIdx.x / p.nv;
    const uint32_t ncg_log = p.row_shift - p.log_c;
    pos0 = ((uint32_t)(tl >> ncg_log) << p.lvl_hi) + ((uint32_t)(tl & ((1ull << ncg_log) - 1)) << p.log_c);
    gbase = (w << p.log_h) + pos0;
  }
  auto goff = [&](uint32_t e) -> unsigned long long {
    return p.packed ? e : ((unsigned long long)(e >> p.log_c) << p.row_shift) + (e & ((1u << p.log_c) - 1));
  };
  const uint32_t hmask = (1u << p.log_h) - 1;
  for (uint32_t e = threadIdx.x; e < T; e += blockDim.x) {
    const unsigned long long g = gbase + goff(e);
    F v = g < p.total ? p.in[g] : 0;
    if (SYM && p.pre) v = fmul(v, __ldg(p.pre + (uint32_t)(g & hmask)));
    tile[e] = v;
  }
  __syncthreads();
  const uint32_t boff = p.packed ? 0u : p.log_c - p.row_shift;   // tile bit of level j is j + boff (mod 2^32)
  for (int phase = 0; phase < 2; phase++) {
    if (phase == 0 ? !p.do_d : !p.do_r) continue;
    for (uint32_t s = 0; s < p.lvl_hi - p.lvl_lo; s++) {
      const uint32_t j = phase == 0 ? p.lvl_hi - 1 - s : p.lvl_lo + s;
      const uint32_t b = j + boff, S = 1u << b;
      const uint4* mats = (phase == 0 ? p.dmat : p.rmat) + (2u << j) + (phase == 0 ? p.dskip : p.rskip);
      const F* tw = (phase == 0 ? p.tw_d : p.tw_r) + (1u << j);
      for (uint32_t q = threadIdx.x; q < T / 2; q += blockDim.x) {
        const uint32_t e0 = ((q >> b) << (b + 1)) | (q & (S - 1)), e1 = e0 + S;
        const uint32_t i = (uint32_t)((pos0 + goff(e0)) & hmask) & ((1u << j) - 1);
        const F x = tile[e0], y = tile[e1];
        if (SYM) {
          const F g = __ldg(tw + i);
          if (phase == 0) {
            tile[e0] = fadd(x, y);
            tile[e1] = fmul(fsub(x, y), g);
          } else {
            const F t = fmul(g, y);
            tile[e0] = fadd(x, t);
            tile[e1] = fsub(x, t);
          }
        } else {
          const uint4 m = __ldg(mats + 2 * i);
          tile[e0] = fdot2(m.x, x, m.y, y);
          tile[e1] = fdot2(m.z, x, m.w, y);
        }
      }
      __syncthreads();
    }
  }
  for (uint32_t e = threadIdx.x; e < T; e += blockDim.x) {
    const unsigned long long g = gbase + goff(e);
    if (g >= p.total) continue;
    F v = tile[e];
    if (SYM && p.post) v = fmul(v, __ldg(p.post + (uint32_t)(g & hmask)));
    p.out[g] = v;
  }
}

// ---------------------------------------------------------------------------------------------------------
// tree
// ---------------------------------------------------------------------------------------------------------
struct Lv {
  uint32_t log_n = 0;
  uint4 *rmat = nullptr, *dmat = nullptr;                                   // N each
  F *xnn = nullptr, *xnn_inv = nullptr, *z0z0 = nullptr, *z1z1 = nullptr;   // N each
  F *z0_s1 = nullptr, *z1_s0 = nullptr, *z0i = nullptr, *z1i = nullptr;     // N/2 each
  // symmetric-butterfly tables (h = N/2 entries each, index = moiety): g and 1/g at entry 2^j + i, the accumulated scale
  // Gamma_p = prod_j (y_j(p) + beta_j) y_j(p)^(2^j - 1) and the pre-scale 2^-L / Gamma_p, and gam[1][i] * xnn[2i+1]
  bool sym = false;
  F *tw_r[2] = {nullptr, nullptr}, *tw_d[2] = {nullptr, nullptr}, *gam[2] = {nullptr, nullptr}, *gami[2] = {nullptr, nullptr};
  F* gx = nullptr;
};
struct Map { F x0, t; };   // r(x) = (x^2 - x0 x + t) / (x - x0), src/ec.rs:231-232

}  // namespace m31
}  // namespace ecfft

struct ecfft_m31_tree {
  int device = 0;
  uint32_t log_n = 0;
  ecfft::m31::F* f = nullptr;                 // 2n, BinaryTree order (f[0] = 0); chain level N reads it with stride n/N
  std::vector<ecfft::m31::Lv> lv;             // lv[k]: 2^k leaves
  std::vector<ecfft::m31::Map> maps;
  ecfft::m31::F leaf2[2] = {0, 0};            // leaves of the 2-leaf chain level (VANISH base case, src/fftree.rs:293-298)
  cudaStream_t st = nullptr;
  std::vector<void*> owned;
  std::mutex mu;
  size_t n() const { return (size_t)1 << log_n; }
  template <class T>
  T* alloc(size_t count) {
    void* p = nullptr;
    ECFFT_CUDA(cudaMalloc(&p, (count ? count : 1) * sizeof(T)));
    owned.push_back(p);
    return (T*)p;
  }
  ~ecfft_m31_tree() {
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(device);
    for (void* p : owned) cudaFree(p);
    if (st) cudaStreamDestroy(st);
    if (prev >= 0 && prev != device) cudaSetDevice(prev);
  }
};

namespace ecfft {
namespace m31 {
typedef ecfft_m31_tree Tree;

static inline bool is_pow2(size_t n) { return n && !(n & (n - 1)); }
static inline uint32_t ilog2(size_t n) {
  uint32_t l = 0;
  while (n >>= 1) l++;
  return l;
}

// The algorithms on device buffers: every recursion depth of the reference is one batched launch set over all
// sub-problems of that depth (they share the chain level's tables), as in engine.cu.
struct Eng {
  const Tree& t;
  cudaStream_t st;
  std::vector<void*> scratch;
  Eng(const Tree& tree, cudaStream_t s) : t(tree), st(s) {}
  ~Eng() {
    for (void* p : scratch) cudaFreeAsync(p, st);
  }
  F* tmp(size_t count) {
    void* p = nullptr;
    ECFFT_CUDA(cudaMallocAsync(&p, (count ? count : 1) * sizeof(F), st));
    scratch.push_back(p);
    return (F*)p;
  }
  const Lv& level_for(size_t leaves) const {   // subtree_with_size, src/fftree.rs:489-496
    if (!is_pow2(leaves)) throw Error(ERR_NOT_POW2, "length is not a power of two");
    const uint32_t lg = ilog2(leaves);
    if (lg > t.log_n) throw Error(ERR_TREE_TOO_SMALL, "FFTree is too small");
    return t.lv[lg];
  }

  void launch(const Pass& p, bool sym) {
    const size_t tiles = (p.total + ((size_t)1 << p.log_t) - 1) >> p.log_t;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)tiles);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = sizeof(F) << p.log_t;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = t_pdl ? 1 : 0;   // see pdl_mode()
    if (sym) ECFFT_CUDA(cudaLaunchKernelEx(&cfg, k31_extend<true>, p));
    else ECFFT_CUDA(cudaLaunchKernelEx(&cfg, k31_extend<false>, p));
    prof::count_launch();
  }
  // EXTEND of nvec contiguous vectors of length h = 2^log_h towards `target` (src/fftree.rs:72-126); in may equal out.
  // unscaled: with the symmetric tables, leave the final Gamma^target scaling to the caller (ENTER's combine carries it).
  void extend(const F* in, F* out, uint32_t log_h, size_t nvec, int target, bool unscaled = false) {
    const Lv& lv = level_for((size_t)2 << log_h);
    const size_t total = nvec << log_h;
    if (log_h == 0) {
      if (in != out) ECFFT_CUDA(cudaMemcpyAsync(out, in, total * sizeof(F), cudaMemcpyDeviceToDevice, st));
      return;
    }
    static const bool force_matrix = getenv("ECFFT_B200_M31_MATRIX") != nullptr;   // A/B switch: the reference's matrix butterflies
    const bool sym = lv.sym && !force_matrix;
    if (unscaled && !sym) throw Error(ERR_INVALID_ARG, "m31 extend: unscaled output needs the symmetric tables");
    Pass p{};
    p.dmat = lv.dmat;
    p.rmat = lv.rmat;
    p.tw_d = sym ? lv.tw_d[1 - target] : nullptr;
    p.tw_r = sym ? lv.tw_r[target] : nullptr;
    const F* pre = sym ? lv.gami[1 - target] : nullptr;
    const F* post = (sym && !unscaled) ? lv.gam[target] : nullptr;
    p.total = total;
    p.log_h = log_h;
    p.dskip = target == 0 ? 1 : 0;   // src/fftree.rs:87-90
    p.rskip = target == 0 ? 0 : 1;   // src/fftree.rs:108-111
    p.log_t = LT;
    if (log_h <= LT) {               // whole vectors per tile: one launch
      p.in = in; p.out = out; p.packed = 1; p.lvl_lo = 0; p.lvl_hi = log_h; p.do_d = p.do_r = 1; p.nv = 1;
      p.pre = pre; p.post = post;
      launch(p, sym);
      return;
    }
    const uint32_t outer = log_h - LT, kmax = LT - 7, npass = (outer + kmax - 1) / kmax;
    std::vector<uint32_t> bounds{log_h};
    for (uint32_t i = 1; i <= npass; i++) bounds.push_back(log_h - (outer * i) / npass);
    const F* src = in;
    p.nv = nvec;
    for (uint32_t i = 0; i < npass; i++) {   // outer decompose passes, top levels first
      p.in = src; p.out = out; p.packed = 0; p.do_d = 1; p.do_r = 0;
      p.lvl_hi = bounds[i]; p.lvl_lo = bounds[i + 1];
      p.krows = p.lvl_hi - p.lvl_lo; p.log_c = LT - p.krows; p.row_shift = p.lvl_lo;
      p.pre = i == 0 ? pre : nullptr; p.post = nullptr;
      launch(p, sym);
      src = out;
    }
    p.in = src; p.out = out; p.packed = 1; p.lvl_lo = 0; p.lvl_hi = LT; p.do_d = p.do_r = 1;
    p.pre = nullptr; p.post = nullptr;
    launch(p, sym);
    for (uint32_t i = npass; i-- > 0;) {     // outer recombine passes, top levels last
      p.in = out; p.out = out; p.packed = 0; p.do_d = 0; p.do_r = 1;
      p.lvl_hi = bounds[i]; p.lvl_lo = bounds[i + 1];
      p.krows = p.lvl_hi - p.lvl_lo; p.log_c = LT - p.krows; p.row_shift = p.lvl_lo;
      p.pre = nullptr; p.post = i == 0 ? post : nullptr;
      launch(p, sym);
    }
  }

  // src/fftree.rs:143-161, bottom-up: after the pass for m the array holds n/m evaluation vectors of length m
  void enter(const F* in, F* out, size_t n) {
    level_for(n);
    if (n == 1) {
      if (in != out) ECFFT_CUDA(cudaMemcpyAsync(out, in, sizeof(F), cudaMemcpyDeviceToDevice, st));
      return;
    }
    F* W = tmp(n);
    F* ping[2] = {tmp(n), tmp(n)};
    const F* cur = in;
    uint32_t idx = 0;
    for (size_t m = 2; m <= n; m *= 2, idx++) {
      const Lv& lv = level_for(m);
      const size_t h = m / 2;
      const uint32_t log_h = ilog2(h);
      F* dst = m == n ? out : ping[idx & 1];
      static const bool force_matrix = getenv("ECFFT_B200_M31_MATRIX") != nullptr;
      const bool sym = lv.sym && !force_matrix && log_h >= 1;
      extend(cur, W, log_h, n / h, 1, sym);
      const F* A = cur;
      const F* xnn = lv.xnn;
      if (sym) {   // the EXTEND left out its Gamma^1 scaling: out[2i+1] = gam1[i] u1^ + (gam1[i] xnn[2i+1]) v1^, one reduction
        const F *gam1 = lv.gam[1], *gx = lv.gx;
        map(n / 2, st, [=] __device__(size_t k) {
          const size_t blk = k >> log_h, i = k & (h - 1), off = blk << (log_h + 1);
          dst[off + 2 * i] = fadd(A[off + i], fmul(A[off + h + i], __ldg(xnn + 2 * i)));
          dst[off + 2 * i + 1] = fdot2(__ldg(gam1 + i), W[off + i], __ldg(gx + i), W[off + h + i]);
        });
      } else {
        map(n / 2, st, [=] __device__(size_t k) {
          const size_t blk = k >> log_h, i = k & (h - 1), off = blk << (log_h + 1);
          dst[off + 2 * i] = fadd(A[off + i], fmul(A[off + h + i], __ldg(xnn + 2 * i)));
          dst[off + 2 * i + 1] = fadd(W[off + i], fmul(W[off + h + i], __ldg(xnn + 2 * i + 1)));
        });
      }
      cur = dst;
    }
  }

  // src/fftree.rs:232-259 for nvec vectors of length len sharing `a`; out may not alias evals
  void redc(const F* evals, const F* a, size_t len, size_t nvec, int moiety, F* out) {
    const Lv& lv = level_for(len);
    if (len < 2) throw Error(ERR_INVALID_ARG, "redc: length must be >= 2");
    const F* zinv = moiety == 0 ? lv.z0i : lv.z1i;
    const size_t h = len / 2;
    const uint32_t log_h = ilog2(h);
    F* t0 = tmp(h * nvec);
    F* g1 = tmp(h * nvec);
    // the reference batch-inverts a[::2] on every call (src/fftree.rs:235); for a = xnn_s the stored inverses are the same values
    if (a == lv.xnn) {
      const F* ai = lv.xnn_inv;
      map(h * nvec, st, [=] __device__(size_t k) { t0[k] = fmul(evals[2 * k], __ldg(ai + 2 * (k & (h - 1)))); });
    } else {
      map(h * nvec, st, [=] __device__(size_t k) { t0[k] = fmul(evals[2 * k], finv(__ldg(a + 2 * (k & (h - 1))))); });
    }
    extend(t0, g1, log_h, nvec, 1 - moiety);
    F* h1 = t0;
    map(h * nvec, st, [=] __device__(size_t k) {
      const size_t i = k & (h - 1);
      h1[k] = fmul(fsub(evals[2 * k + 1], fmul(g1[k], __ldg(a + 2 * i + 1))), __ldg(zinv + i));
    });
    F* h0 = g1;
    extend(h1, h0, log_h, nvec, moiety);
    map(h * nvec, st, [=] __device__(size_t k) {
      out[2 * k] = h0[k];
      out[2 * k + 1] = h1[k];
    });
  }
  // src/fftree.rs:277-281
  void mod(const F* evals, const F* a, const F* c, size_t len, size_t nvec, F* out) {
    F* hb = tmp(len * nvec);
    redc(evals, a, len, nvec, 0, hb);
    map(len * nvec, st, [=] __device__(size_t k) { hb[k] = fmul(hb[k], __ldg(c + (k & (len - 1)))); });
    redc(hb, a, len, nvec, 0, out);
  }
  // src/fftree.rs:200-224, top-down: before the pass for m the array holds n/m evaluation vectors of length m
  void exit(const F* evals, F* out, size_t n) {
    level_for(n);
    F* cur = tmp(n);
    F* nxt = tmp(n);
    F* M = tmp(n);
    ECFFT_CUDA(cudaMemcpyAsync(cur, evals, n * sizeof(F), cudaMemcpyDeviceToDevice, st));
    for (size_t m = n; m >= 2; m /= 2) {
      const Lv& lv = level_for(m);
      const size_t h = m / 2;
      const uint32_t log_h = ilog2(h);
      mod(cur, lv.xnn, lv.z0z0, m, n / m, M);
      const F* xinv = lv.xnn_inv;
      const F *c = cur, *Mm = M;
      F* nx = nxt;
      map(n / 2, st, [=] __device__(size_t k) {
        const size_t v = k >> log_h, i = k & (h - 1), off = v << (log_h + 1);
        const F u0 = Mm[off + 2 * i];
        nx[off + i] = u0;
        nx[off + h + i] = fmul(fsub(c[off + 2 * i], u0), __ldg(xinv + 2 * i));
      });
      std::swap(cur, nxt);
    }
    ECFFT_CUDA(cudaMemcpyAsync(out, cur, n * sizeof(F), cudaMemcpyDeviceToDevice, st));
  }
  // src/fftree.rs:128-135
  void mextend(const F* in, F* out, size_t h, int target) {
    const Lv& lv = level_for(2 * h);
    const F* z = target == 1 ? lv.z0_s1 : lv.z1_s0;
    extend(in, out, ilog2(h), 1, target);
    map(h, st, [=] __device__(size_t i) { out[i] = fadd(out[i], __ldg(z + i)); });
  }
  // src/fftree.rs:169-192: one data-dependent branch per level, read back as a count of differing positions
  size_t degree(const F* evals, size_t n) {
    level_for(n);
    if (n == 1) return 0;
    F* e0 = tmp(n / 2);
    F* e1 = tmp(n / 2);
    F* g1 = tmp(n / 2);
    F* curbuf = tmp(n);
    unsigned long long* counter = (unsigned long long*)tmp(2);
    const F* cur = evals;
    size_t result = 0;
    for (size_t len = n; len > 1; len /= 2) {
      const Lv& lv = level_for(len);
      const size_t h = len / 2;
      const F* c = cur;
      map(h, st, [=] __device__(size_t i) {
        e0[i] = c[2 * i];
        e1[i] = c[2 * i + 1];
      });
      extend(e0, g1, ilog2(h), 1, 1);
      ECFFT_CUDA(cudaMemsetAsync(counter, 0, sizeof(unsigned long long), st));
      map(h, st, [=] __device__(size_t i) {
        if (g1[i] != e1[i]) atomicAdd(counter, 1ull);
      });
      unsigned long long diff = 0;
      ECFFT_CUDA(cudaMemcpyAsync(&diff, counter, sizeof diff, cudaMemcpyDeviceToHost, st));
      ECFFT_CUDA(cudaStreamSynchronize(st));
      if (diff == 0) {
        ECFFT_CUDA(cudaMemcpyAsync(curbuf, e0, h * sizeof(F), cudaMemcpyDeviceToDevice, st));
      } else {
        const F* zi = lv.z0i;
        map(h, st, [=] __device__(size_t i) { e1[i] = fmul(fsub(e1[i], g1[i]), __ldg(zi + i)); });
        extend(e1, curbuf, ilog2(h), 1, 0);
        result += h;
      }
      cur = curbuf;
    }
    return result;
  }
  // src/fftree.rs:291-308, bottom-up; out has 2n elements
  void vanish(const F* dom, F* out, size_t n) {
    level_for(2 * n);
    F* Q = tmp(2 * n);
    F* Q2 = tmp(2 * n);
    F* q0 = tmp(n);
    F* e = tmp(n);
    const F l0 = t.leaf2[0], l1 = t.leaf2[1];
    map(n, st, [=] __device__(size_t i) {
      Q[2 * i] = fsub(dom[i], l0);
      Q[2 * i + 1] = fsub(dom[i], l1);
    });
    for (size_t len = 2, cnt = n; cnt > 1; len *= 2, cnt /= 2) {
      const Lv& lv = level_for(2 * len);
      const size_t pairs = cnt / 2;
      const uint32_t log_len = ilog2(len);
      const F* Qc = Q;
      map(len * pairs, st, [=] __device__(size_t k) {
        const size_t w = k >> log_len, i = k & (len - 1);
        q0[k] = fmul(Qc[(2 * w) * len + i], Qc[(2 * w + 1) * len + i]);
      });
      extend(q0, e, log_len, pairs, 1);
      const F* z = lv.z0_s1;
      F* Qn = Q2;
      map(len * pairs, st, [=] __device__(size_t k) {
        const size_t w = k >> log_len, i = k & (len - 1);
        Qn[w * 2 * len + 2 * i] = q0[k];
        Qn[w * 2 * len + 2 * i + 1] = fadd(e[k], __ldg(z + i));
      });
      std::swap(Q, Q2);
    }
    ECFFT_CUDA(cudaMemcpyAsync(out, Q, 2 * n * sizeof(F), cudaMemcpyDeviceToDevice, st));
  }
};

// ---------------------------------------------------------------------------------------------------------
// construction: build_ec_fftree (src/ec.rs:498-554) — the O(log n) isogeny chain on the host, everything of size n
// on the device
// ---------------------------------------------------------------------------------------------------------
struct Pt { F x, y; bool inf; };
static Pt padd(Pt p1, Pt p2, F a, F b) {   // src/ec.rs:376-424 with a1 = a2 = a3 = 0
  if (p1.inf) return p2;
  if (p2.inf) return p1;
  if (p1.x == p2.x && fadd(p1.y, p2.y) == 0) return Pt{0, 0, true};
  F lam, nu;
  if (p1.x == p2.x) {
    const F xx = fmul(p1.x, p1.x), d = finv(fadd(p1.y, p1.y));
    lam = fmul(fadd(fadd(fadd(xx, xx), xx), a), d);
    nu = fmul(fadd(fadd(fsub(fmul(a, p1.x), fmul(xx, p1.x)), b), b), d);
  } else {
    const F d = finv(fsub(p2.x, p1.x));
    lam = fmul(fsub(p2.y, p1.y), d);
    nu = fmul(fsub(fmul(p1.y, p2.x), fmul(p2.y, p1.x)), d);
  }
  const F x3 = fsub(fsub(fmul(lam, lam), p1.x), p2.x);
  return Pt{x3, fsub(fneg(fmul(lam, x3)), nu), false};
}
static int two_adicity(Pt p, F a, F b) {   // src/utils.rs:356-365
  for (int i = 0; i < 2048; i++) {
    if (p.inf) return i;
    p = padd(p, p, a, b);
  }
  return -1;
}
// polynomials of degree < 3 modulo the monic cubic x^3 + a x + b
typedef std::array<F, 3> Q3;
static Q3 q3mul(const Q3& u, const Q3& v, F a, F b) {
  F c[5] = {0, 0, 0, 0, 0};
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) c[i + j] = fadd(c[i + j], fmul(u[i], v[j]));
  for (int k = 4; k >= 3; k--) {   // x^k = -(a x^(k-2) + b x^(k-3))
    c[k - 2] = fsub(c[k - 2], fmul(c[k], a));
    c[k - 3] = fsub(c[k - 3], fmul(c[k], b));
  }
  return Q3{c[0], c[1], c[2]};
}
static Q3 q3pow(Q3 base, uint64_t e, F a, F b) {
  Q3 r{1, 0, 0};
  while (e) {
    if (e & 1) r = q3mul(r, base, a, b);
    base = q3mul(base, base, a, b);
    e >>= 1;
  }
  return r;
}
static std::vector<F> poly_gcd(std::vector<F> f, std::vector<F> g) {
  auto trim = [](std::vector<F>& v) { while (!v.empty() && v.back() == 0) v.pop_back(); };
  trim(f); trim(g);
  while (!g.empty()) {
    while (f.size() >= g.size()) {
      const F c = fmul(f.back(), finv(g.back()));
      const size_t s = f.size() - g.size();
      for (size_t i = 0; i < g.size(); i++) f[s + i] = fsub(f[s + i], fmul(c, g[i]));
      trim(f);
      if (f.empty()) break;
    }
    std::swap(f, g);
  }
  if (!f.empty()) {
    const F li = finv(f.back());
    for (F& c : f) c = fmul(c, li);
  }
  return f;
}
// roots of x^3 + a x + b in F_p: the x-coordinates of the 2-torsion points (src/ec.rs:246-260; the reference calls
// its generic find_roots, src/utils.rs:25-226).  gcd with x^p - x, then equal-degree splitting.
static std::vector<F> cubic_roots(F a, F b) {
  std::vector<F> f{b, a, 0, 1};
  Q3 xp = q3pow(Q3{0, 1, 0}, P31, a, b);
  std::vector<F> d{xp[0], fsub(xp[1], 1), xp[2]};
  std::vector<F> g = poly_gcd(f, d);
  if (g.empty()) g = f;
  std::vector<F> roots;
  std::vector<std::vector<F>> stack{g};
  F s = 1;
  while (!stack.empty()) {
    std::vector<F> h = stack.back();
    stack.pop_back();
    if (h.size() <= 1) continue;
    if (h.size() == 2) { roots.push_back(fmul(fneg(h[0]), finv(h[1]))); continue; }
    for (;; s++) {
      // (x + s)^((p-1)/2) - 1 modulo the cubic (h divides it), then gcd with h
      Q3 w = q3pow(Q3{s, 1, 0}, (P31 - 1) / 2, a, b);
      std::vector<F> wv{fsub(w[0], 1), w[1], w[2]};
      std::vector<F> dd = poly_gcd(h, wv);
      if (dd.size() > 1 && dd.size() < h.size()) {
        std::vector<F> q(h.size() - dd.size() + 1, 0), rem = h;   // h / dd, exact
        while (rem.size() >= dd.size()) {
          const F c = rem.back();   // dd is monic
          const size_t sh = rem.size() - dd.size();
          q[sh] = c;
          for (size_t i = 0; i < dd.size(); i++) rem[sh + i] = fsub(rem[sh + i], fmul(c, dd[i]));
          while (!rem.empty() && rem.back() == 0) rem.pop_back();
        }
        stack.push_back(dd);
        stack.push_back(q);
        s++;
        break;
      }
    }
  }
  return roots;
}

// per-thread double-and-add: leaf i = x(offset + i G), src/ec.rs:545-551
__device__ inline void dev_padd(F& x1, F& y1, bool& inf1, F x2, F y2, F a, F b) {
  if (inf1) { x1 = x2; y1 = y2; inf1 = false; return; }
  if (x1 == x2 && fadd(y1, y2) == 0) { inf1 = true; return; }
  F lam, nu;
  if (x1 == x2) {
    const F xx = fmul(x1, x1), d = finv(fadd(y1, y1));
    lam = fmul(fadd(fadd(fadd(xx, xx), xx), a), d);
    nu = fmul(fadd(fadd(fsub(fmul(a, x1), fmul(xx, x1)), b), b), d);
  } else {
    const F d = finv(fsub(x2, x1));
    lam = fmul(fsub(y2, y1), d);
    nu = fmul(fsub(fmul(y1, x2), fmul(y2, x1)), d);
  }
  const F x3 = fsub(fsub(fmul(lam, lam), x1), x2);
  y1 = fsub(fneg(fmul(lam, x3)), nu);
  x1 = x3;
}

static void build_level(Tree& t, uint32_t k, Eng& eng);

static Tree* build(size_t n, int device) {
  // src/lib.rs:199-206
  const F ca = 1, cb = 0;
  const Pt offset{1048755163u, 279503108u, false};
  Pt g{1273083559u, 804329170u, false};
  const uint32_t two_adic = 28;
  if (!is_pow2(n)) throw Error(ERR_NOT_POW2, "n is not a power of two");
  t_pdl = false;
  const uint32_t log_n = ilog2(n);
  if (log_n > two_adic) throw Error(ERR_TOO_LARGE, "FFTree size is too large for the generator (log2 n > 28)");   // src/ec.rs:513-515
  for (uint32_t i = 0; i < two_adic - log_n; i++) g = padd(g, g, ca, cb);
  const Pt generator = g;
  std::unique_ptr<Tree> t(new Tree());
  t->device = device;
  t->log_n = log_n;
  {   // scratch comes from the stream-ordered pool: keep freed blocks instead of returning them to the driver at every sync
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      uint64_t thr = UINT64_MAX;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
  }
  ECFFT_CUDA(cudaStreamCreateWithFlags(&t->st, cudaStreamNonBlocking));
  // the chain of 2-isogenies that each lower the generator's order (src/ec.rs:523-543)
  F a = ca, b = cb;
  for (uint32_t i = 0; i < log_n; i++) {
    const int kk = two_adicity(g, a, b);
    bool found = false;
    for (F x0 : cubic_roots(a, b)) {
      const F tt = fadd(fmul(3, fmul(x0, x0)), a);
      const F a2 = fsub(a, fmul(5, tt)), b2 = fsub(b, fmul(7, fmul(x0, tt)));
      // phi(x, y) = ((x^2 - x0 x + t)/(x - x0), ((x - x0)^2 - t)/(x - x0)^2 y), src/ec.rs:230-236
      Pt gp{0, 0, true};
      const F dx = fsub(g.x, x0);
      if (!g.inf && dx != 0) {
        const F di = finv(dx), d2 = fmul(dx, dx);
        gp = Pt{fmul(fadd(fsub(fmul(g.x, g.x), fmul(x0, g.x)), tt), di), fmul(fmul(fsub(d2, tt), finv(d2)), g.y), false};
      }
      const int kp = two_adicity(gp, a2, b2);
      if (kk >= 0 && kp >= 0 && kk == kp + 1) {
        g = gp;
        a = a2;
        b = b2;
        t->maps.push_back(Map{x0, tt});
        found = true;
        break;
      }
    }
    if (!found) throw Error(ERR_INVALID_ARG, "cannot find a suitable isogeny");   // src/ec.rs:541
  }
  // leaves and internal nodes (FFTree::new, src/fftree.rs:42-67)
  t->f = t->alloc<F>(2 * n);
  ECFFT_CUDA(cudaMemsetAsync(t->f, 0, sizeof(F), t->st));
  std::vector<F> gt(2 * (log_n ? log_n : 1));
  {
    Pt q = generator;
    for (uint32_t j = 0; j < log_n; j++) {
      gt[2 * j] = q.x;
      gt[2 * j + 1] = q.y;
      q = padd(q, q, ca, cb);
    }
  }
  F* gtab = t->alloc<F>(gt.size());
  ECFFT_CUDA(cudaMemcpyAsync(gtab, gt.data(), gt.size() * sizeof(F), cudaMemcpyHostToDevice, t->st));
  ECFFT_CUDA(cudaStreamSynchronize(t->st));
  {
    F* leaves = t->f + n;
    const F ox = offset.x, oy = offset.y;
    map(n, t->st, [=] __device__(size_t i) {
      F x = ox, y = oy;
      bool inf = false;
      for (uint32_t j = 0; j < log_n; j++)
        if ((i >> j) & 1) dev_padd(x, y, inf, gtab[2 * j], gtab[2 * j + 1], ca, cb);
      leaves[i] = x;
    });
  }
  unsigned long long* err = (unsigned long long*)t->alloc<unsigned long long>(1);
  ECFFT_CUDA(cudaMemsetAsync(err, 0, sizeof(unsigned long long), t->st));
  for (uint32_t i = 0; i < log_n; i++) {
    const size_t size = n >> (i + 1);
    const F* prev = t->f + 2 * size;
    F* layer = t->f + size;
    const F x0 = t->maps[i].x0, tt = t->maps[i].t;
    map(size, t->st, [=] __device__(size_t k) {
      F v[2];
      for (int s = 0; s < 2; s++) {
        const F x = prev[k + s * size], d = fsub(x, x0);
        if (d == 0) atomicAdd(err, 1ull);   // the reference unwraps: rational_map.map(..).unwrap()
        v[s] = fmul(fadd(fsub(fmul(x, x), fmul(x0, x)), tt), finv(d));
      }
      if (v[0] != v[1]) atomicAdd(err, 1ull);   // debug_assert_eq, src/fftree.rs:64
      layer[k] = v[0];
    });
  }
  t->lv.resize(log_n + 1);
  Eng eng(*t, t->st);
  for (uint32_t k = 0; k <= log_n; k++) build_level(*t, k, eng);
  unsigned long long bad = 0;
  ECFFT_CUDA(cudaMemcpyAsync(&bad, err, sizeof bad, cudaMemcpyDeviceToHost, t->st));
  ECFFT_CUDA(cudaStreamSynchronize(t->st));
  if (bad) throw Error(ERR_INVALID_ARG, "tree construction met a zero denominator or an inconsistent rational map");
  return t.release();
}

// Symmetric-butterfly tables of the chain level with N = 2^k leaves: the level with half-stride 2^j
// uses the level's f layer with 2^(j+2) nodes and that layer's map (x0_j, t_j = beta_j^2).
static void build_sym_tables(Tree& t, uint32_t k) {
  Lv& lv = t.lv[k];
  if (k < 2) return;
  const uint32_t L = k - 1;
  const size_t n = t.n(), N = (size_t)1 << k, stride = n / N, h = N / 2;
  std::vector<F> host(2 * L);
  for (uint32_t j = 0; j < L; j++) {
    const Map& m = t.maps[k - 2 - j];
    const F beta = fpow(m.t, ((uint64_t)P31 + 1) / 4);   // p = 3 (mod 4)
    if (fmul(beta, beta) != m.t) return;                  // t is not a square: keep the matrix butterflies for this level
    host[2 * j] = m.x0;
    host[2 * j + 1] = beta;
  }
  cudaStream_t st = t.st;
  F* xb = t.alloc<F>(2 * L);
  ECFFT_CUDA(cudaMemcpyAsync(xb, host.data(), 2 * L * sizeof(F), cudaMemcpyHostToDevice, st));
  ECFFT_CUDA(cudaStreamSynchronize(st));
  const F* f = t.f;
  F inv2L = 1;
  for (uint32_t j = 0; j < L; j++) inv2L = fmul(inv2L, (P31 + 1) / 2);
  for (int mu = 0; mu < 2; mu++) {
    F *twr = lv.tw_r[mu] = t.alloc<F>(h), *twd = lv.tw_d[mu] = t.alloc<F>(h);
    F *gam = lv.gam[mu] = t.alloc<F>(h), *gami = lv.gami[mu] = t.alloc<F>(h);
    map(h, st, [=] __device__(size_t idx) {
      if (idx == 0) { twr[0] = twd[0] = 0; return; }
      const uint32_t j = 31 - __clz((unsigned)idx);
      const size_t i = idx - ((size_t)1 << j);
      const F* layer = f + (n >> (k - 2 - j));
      const F y0 = fsub(layer[(2 * i + mu) * stride], xb[2 * j]), beta = xb[2 * j + 1];
      const F num = fsub(y0, beta), den = fadd(y0, beta);
      twr[idx] = fmul(num, finv(den));
      twd[idx] = fmul(den, finv(num));
    });
    map(h, st, [=] __device__(size_t pp) {
      F acc = 1;
      for (uint32_t j = 0; j < L; j++) {
        const size_t i = pp & (((size_t)1 << j) - 1), bit = (pp >> j) & 1;
        const F* layer = f + (n >> (k - 2 - j));
        const F y = fsub(layer[(2 * i + mu + bit * ((size_t)2 << j)) * stride], xb[2 * j]);
        acc = fmul(acc, fmul(fadd(y, xb[2 * j + 1]), fpow(y, ((uint64_t)1 << j) - 1)));
      }
      gam[pp] = acc;
      gami[pp] = fmul(inv2L, finv(acc));
    });
  }
  {
    F* gx = lv.gx = t.alloc<F>(h);
    const F *gam1 = lv.gam[1], *xnn = lv.xnn;
    map(h, st, [=] __device__(size_t i) { gx[i] = fmul(gam1[i], xnn[2 * i + 1]); });
  }
  lv.sym = true;
}

// from_tree (src/fftree.rs:318-463) for the chain level with N = 2^k leaves = every (n/N)-th leaf of the top tree;
// the levels below are complete (the reference derives the subtree first, :319).
static void build_level(Tree& t, uint32_t k, Eng& eng) {
  const size_t n = t.n(), N = (size_t)1 << k, stride = n / N, h = N / 2;
  cudaStream_t st = t.st;
  Lv& lv = t.lv[k];
  lv.log_n = k;
  const F* f = t.f;
  // matrices, src/fftree.rs:340-363: layer with d pairs uses map (index of that layer) and exponent d/2 - 1
  lv.rmat = t.alloc<uint4>(N);
  lv.dmat = t.alloc<uint4>(N);
  {
    uint4 *R = lv.rmat, *D = lv.dmat;
    map(N, st, [=] __device__(size_t i) { R[i] = D[i] = make_uint4(1, 0, 0, 1); });
    for (uint32_t li = 0; li < k; li++) {
      const size_t d = N >> (li + 1);   // layer li of this level's f has 2d nodes at f[(2d) * stride ...] of the top tree
      if (d == 1) continue;
      const F x0 = t.maps[li].x0;
      const F* layer = f + (n >> li);   // the top tree's layer li; this level reads it with `stride`
      map(d, st, [=] __device__(size_t i) {
        const F s0 = layer[i * stride], s1 = layer[(i + d) * stride];
        const F v0 = fpow(fsub(s0, x0), d / 2 - 1), v1 = fpow(fsub(s1, x0), d / 2 - 1);
        const uint4 r = make_uint4(v0, fmul(s0, v0), v1, fmul(s1, v1));
        const F di = finv(fsub(fmul(r.x, r.w), fmul(r.y, r.z)));
        R[d + i] = r;
        D[d + i] = make_uint4(fmul(r.w, di), fmul(fneg(r.y), di), fmul(fneg(r.z), di), fmul(r.x, di));
      });
    }
  }
  lv.xnn = t.alloc<F>(N);
  lv.xnn_inv = t.alloc<F>(N);
  F* xq = eng.tmp(N);
  F* xq_inv = eng.tmp(N);
  {
    F *xn = lv.xnn, *xi = lv.xnn_inv;
    const F* leaves = f + n;
    map(N, st, [=] __device__(size_t i) {
      const F x = leaves[i * stride];
      const F q = fpow(x, N / 4), v = fpow(x, N / 2);
      xq[i] = q;
      xq_inv[i] = finv(q);
      xn[i] = v;
      xi[i] = finv(v);
    });
  }
  if (k == 0) return;
  build_sym_tables(t, k);   // from here on this level's EXTENDs run the one-product butterflies
  lv.z0_s1 = t.alloc<F>(h);
  lv.z1_s0 = t.alloc<F>(h);
  lv.z0i = t.alloc<F>(h);
  lv.z1i = t.alloc<F>(h);
  lv.z0z0 = t.alloc<F>(N);
  lv.z1z1 = t.alloc<F>(N);
  const F* leaves = f + n;
  if (k == 1) {   // base cases, src/fftree.rs:400-403, 455-458
    F *a = lv.z0_s1, *b = lv.z1_s0, *zz0 = lv.z0z0, *zz1 = lv.z1z1;
    map(1, st, [=] __device__(size_t) {
      const F s0 = leaves[0], s1 = leaves[stride];
      a[0] = fsub(s1, s0);
      b[0] = fsub(s0, s1);
      zz0[0] = zz0[1] = fmul(s0, s0);
      zz1[0] = zz1[1] = fmul(s1, s1);
    });
    ECFFT_CUDA(cudaMemcpyAsync(&t.leaf2[0], leaves, sizeof(F), cudaMemcpyDeviceToHost, st));
    ECFFT_CUDA(cudaMemcpyAsync(&t.leaf2[1], leaves + stride, sizeof(F), cudaMemcpyDeviceToHost, st));
    ECFFT_CUDA(cudaStreamSynchronize(st));
  } else {
    const Lv& sub = t.lv[k - 1];
    // z0_s1 from the subtree's vanishing polynomials, src/fftree.rs:383-392
    F* u = eng.tmp(h);
    F* v = eng.tmp(h);
    {
      const F *sz0 = sub.z0_s1, *sz1 = sub.z1_s0;
      map(h / 2, st, [=] __device__(size_t i) {
        u[2 * i] = 0; u[2 * i + 1] = sz0[i];
        v[2 * i] = sz1[i]; v[2 * i + 1] = 0;
      });
    }
    eng.extend(u, u, k - 1, 1, 1);
    eng.extend(v, v, k - 1, 1, 1);
    {
      F* z = lv.z0_s1;
      map(h, st, [=] __device__(size_t i) { z[i] = fmul(u[i], v[i]); });
    }
    // z1_s0 = vanish(S1)[::2], src/fftree.rs:394-396 (vanish uses z0_s1 of this level and the levels below)
    F* s1 = eng.tmp(h);
    F* z1s = eng.tmp(N);
    map(h, st, [=] __device__(size_t i) { s1[i] = leaves[(2 * i + 1) * stride]; });
    eng.vanish(s1, z1s, h);
    {
      F* z = lv.z1_s0;
      map(h, st, [=] __device__(size_t i) { z[i] = z1s[2 * i]; });
    }
  }
  {
    F *a = lv.z0i, *b = lv.z1i;
    const F *z0 = lv.z0_s1, *z1 = lv.z1_s0;
    map(h, st, [=] __device__(size_t i) {
      a[i] = finv(z0[i]);
      b[i] = finv(z1[i]);
    });
  }
  if (k >= 2) {   // src/fftree.rs:418-453
    const Lv& sub = t.lv[k - 1];
    F* sq = eng.tmp(h);
    F* e0 = eng.tmp(h);
    F* e1 = eng.tmp(h);
    F* zz4 = eng.tmp(N);
    F* q = eng.tmp(N);
    F* hi = eng.tmp(N);
    {
      const F *a = sub.z0z0, *b = sub.z1z1;
      map(h, st, [=] __device__(size_t i) { sq[i] = fmul(a[i], b[i]); });
    }
    eng.mod(sq, sub.xnn, sub.z0z0, h, 1, e0);
    eng.extend(e0, e1, k - 1, 1, 1);
    {
      const F *z0 = lv.z0_s1, *xn = lv.xnn;
      map(h, st, [=] __device__(size_t i) {
        zz4[2 * i] = e0[i];
        zz4[2 * i + 1] = e1[i];
      });
      map(N, st, [=] __device__(size_t i) {
        const F z = (i & 1) ? z0[i >> 1] : 0;
        const F r = fsub(z, xn[i]);
        q[i] = fmul(fsub(fmul(r, r), zz4[i]), xq_inv[i]);
      });
    }
    eng.mod(q, xq, zz4, N, 1, hi);
    {
      F *zz0 = lv.z0z0, *zz1 = lv.z1z1;
      const F *z1 = lv.z1_s0, *xn = lv.xnn;
      map(N, st, [=] __device__(size_t i) { zz0[i] = fadd(zz4[i], fmul(xq[i], hi[i])); });
      map(N, st, [=] __device__(size_t i) {
        const F z = (i & 1) ? 0 : z1[i >> 1];
        const F r = fsub(z, xn[i]);
        q[i] = fmul(r, r);
      });
      eng.mod(q, lv.xnn, lv.z0z0, N, 1, zz1);
    }
  }
}

}  // namespace m31
}  // namespace ecfft

// ---------------------------------------------------------------------------------------------------------
// C ABI (include/ecfft_b200.h, "m31")
// ---------------------------------------------------------------------------------------------------------
using namespace ecfft;
namespace {
template <class Fn>
int guard31(Fn fn) {
  try {
    fn();
    return ECFFT_OK;
  } catch (const Error& e) {
    set_last_error(e.what());
    return e.code;
  } catch (const std::exception& e) {
    set_last_error(e.what());
    return ECFFT_ERR_INVALID_ARG;
  }
}
void need(bool ok, int code, const char* msg) {
  if (!ok) throw Error(code, msg);
}
// host-buffer call: upload the operands, run fn(engine, device operands..., device out), download
struct Io {
  DeviceGuard dev;
  m31::Tree& t;
  m31::Eng eng;
  Io(const ecfft_m31_tree* tree) : dev(tree->device), t(*const_cast<ecfft_m31_tree*>(tree)), eng(t, t.st) {}
  m31::F* in(const uint32_t* host, size_t n) {
    need(host != nullptr || n == 0, ERR_INVALID_ARG, "null input buffer");
    m31::F* d = eng.tmp(n);
    if (n) ECFFT_CUDA(cudaMemcpyAsync(d, host, n * sizeof(m31::F), cudaMemcpyHostToDevice, t.st));
    return d;
  }
  void out(uint32_t* host, const m31::F* d, size_t n) {
    need(host != nullptr || n == 0, ERR_INVALID_ARG, "null output buffer");
    if (n) ECFFT_CUDA(cudaMemcpyAsync(host, d, n * sizeof(m31::F), cudaMemcpyDeviceToHost, t.st));
    ECFFT_CUDA(cudaStreamSynchronize(t.st));
  }
};
}  // namespace

#define M31_LOCKED                                                   \
  need(t != nullptr, ERR_INVALID_ARG, "null tree handle");           \
  std::lock_guard<std::mutex> lock(const_cast<ecfft_m31_tree*>(t)->mu); \
  m31::set_call_size(n);                                             \
  Io io(t);

extern "C" {

int ecfft_m31_tree_build(size_t n, int device, ecfft_m31_tree** out) {
  return guard31([&] {
    need(out != nullptr, ERR_INVALID_ARG, "null out");
    need(n && !(n & (n - 1)), ERR_NOT_POW2, "n is not a power of two");
    need(n <= ((size_t)1 << 28), ERR_TOO_LARGE, "FFTree size is too large for the generator (log2 n > 28)");
    DeviceGuard g(device);
    *out = m31::build(n, device);
  });
}
void ecfft_m31_tree_free(ecfft_m31_tree* t) { delete t; }
size_t ecfft_m31_tree_leaves(const ecfft_m31_tree* t) { return t ? t->n() : 0; }

int ecfft_m31_tree_table(const ecfft_m31_tree* t, size_t subtree_leaves, const char* name, uint32_t* out, size_t cap, size_t* count) {
  return guard31([&] {
    need(t && name && count, ERR_INVALID_ARG, "null argument");
    const size_t n = subtree_leaves;
    M31_LOCKED
    const m31::Lv& lv = io.eng.level_for(subtree_leaves);
    const size_t N = subtree_leaves, h = N / 2, n_top = t->n();
    const std::string nm(name);
    const void* src = nullptr;
    size_t cnt = 0;
    m31::F* staged = nullptr;
    if (nm == "f") {
      cnt = 2 * N;
      staged = io.eng.tmp(cnt);
      ECFFT_CUDA(cudaMemsetAsync(staged, 0, sizeof(m31::F), t->st));
      const m31::F* f = t->f;
      const size_t stride = n_top / N;
      for (uint32_t kk = 0; kk <= lv.log_n; kk++) {
        m31::F* dst = staged + (N >> kk);
        const m31::F* s = f + (n_top >> kk);
        m31::map(N >> kk, t->st, [=] __device__(size_t i) { dst[i] = s[i * stride]; });
      }
      src = staged;
    } else if (nm == "recombine_matrices") { src = lv.rmat; cnt = 4 * N; }
    else if (nm == "decompose_matrices") { src = lv.dmat; cnt = 4 * N; }
    else if (nm == "xnn_s") { src = lv.xnn; cnt = N; }
    else if (nm == "xnn_s_inv") { src = lv.xnn_inv; cnt = N; }
    else if (nm == "z0_s1") { src = lv.z0_s1; cnt = h; }
    else if (nm == "z1_s0") { src = lv.z1_s0; cnt = h; }
    else if (nm == "z0_inv_s1") { src = lv.z0i; cnt = h; }
    else if (nm == "z1_inv_s0") { src = lv.z1i; cnt = h; }
    else if (nm == "z0z0_rem_xnn_s") { src = lv.z0z0; cnt = N > 1 ? N : 0; }
    else if (nm == "z1z1_rem_xnn_s") { src = lv.z1z1; cnt = N > 1 ? N : 0; }
    else throw Error(ERR_INVALID_ARG, "unknown table name");
    *count = cnt;
    if (!out) return;
    need(cap >= cnt, ERR_BUFFER_TOO_SMALL, "table buffer too small");
    io.out(out, (const m31::F*)src, cnt);
  });
}

int ecfft_m31_enter(const ecfft_m31_tree* t, const uint32_t* coeffs, size_t n, uint32_t* evals) {
  return guard31([&] {
    M31_LOCKED
    io.eng.level_for(n);
    m31::F* d = io.in(coeffs, n);
    m31::F* o = io.eng.tmp(n);
    io.eng.enter(d, o, n);
    io.out(evals, o, n);
  });
}
int ecfft_m31_exit(const ecfft_m31_tree* t, const uint32_t* evals, size_t n, uint32_t* coeffs) {
  return guard31([&] {
    M31_LOCKED
    io.eng.level_for(n);
    m31::F* d = io.in(evals, n);
    m31::F* o = io.eng.tmp(n);
    io.eng.exit(d, o, n);
    io.out(coeffs, o, n);
  });
}
int ecfft_m31_extend(const ecfft_m31_tree* t, const uint32_t* evals, size_t n, int moiety, uint32_t* out) {
  return guard31([&] {
    M31_LOCKED
    need(moiety == 0 || moiety == 1, ERR_INVALID_ARG, "bad moiety");
    need(n > 0 && n <= ((size_t)1 << 30), ERR_NOT_POW2, "bad length");
    io.eng.level_for(2 * n);
    m31::F* d = io.in(evals, n);
    m31::F* o = io.eng.tmp(n);
    io.eng.extend(d, o, m31::ilog2(n), 1, moiety);
    io.out(out, o, n);
  });
}
int ecfft_m31_mextend(const ecfft_m31_tree* t, const uint32_t* evals, size_t n, int moiety, uint32_t* out) {
  return guard31([&] {
    M31_LOCKED
    need(moiety == 0 || moiety == 1, ERR_INVALID_ARG, "bad moiety");
    need(n > 0 && n <= ((size_t)1 << 30), ERR_NOT_POW2, "bad length");
    io.eng.level_for(2 * n);
    m31::F* d = io.in(evals, n);
    m31::F* o = io.eng.tmp(n);
    io.eng.mextend(d, o, n, moiety);
    io.out(out, o, n);
  });
}
int ecfft_m31_degree(const ecfft_m31_tree* t, const uint32_t* evals, size_t n, size_t* degree) {
  return guard31([&] {
    M31_LOCKED
    need(degree != nullptr, ERR_INVALID_ARG, "null degree");
    io.eng.level_for(n);
    m31::F* d = io.in(evals, n);
    *degree = io.eng.degree(d, n);
  });
}
static int m31_redc(const ecfft_m31_tree* t, const uint32_t* evals, const uint32_t* a, size_t n, int moiety, uint32_t* out) {
  return guard31([&] {
    M31_LOCKED
    io.eng.level_for(n);
    need(n >= 2, ERR_INVALID_ARG, "redc: length must be >= 2");
    m31::F* d = io.in(evals, n);
    m31::F* da = io.in(a, n);
    m31::F* o = io.eng.tmp(n);
    io.eng.redc(d, da, n, 1, moiety, o);
    io.out(out, o, n);
  });
}
int ecfft_m31_redc_z0(const ecfft_m31_tree* t, const uint32_t* evals, const uint32_t* a, size_t n, uint32_t* out) { return m31_redc(t, evals, a, n, 0, out); }
int ecfft_m31_redc_z1(const ecfft_m31_tree* t, const uint32_t* evals, const uint32_t* a, size_t n, uint32_t* out) { return m31_redc(t, evals, a, n, 1, out); }
int ecfft_m31_modular_reduce(const ecfft_m31_tree* t, const uint32_t* evals, const uint32_t* a, const uint32_t* c, size_t n, uint32_t* out) {
  return guard31([&] {
    M31_LOCKED
    io.eng.level_for(n);
    need(n >= 2, ERR_INVALID_ARG, "modular_reduce: length must be >= 2");
    m31::F* d = io.in(evals, n);
    m31::F* da = io.in(a, n);
    m31::F* dc = io.in(c, n);
    m31::F* o = io.eng.tmp(n);
    io.eng.mod(d, da, dc, n, 1, o);
    io.out(out, o, n);
  });
}
int ecfft_m31_vanish(const ecfft_m31_tree* t, const uint32_t* domain, size_t n, uint32_t* out) {
  return guard31([&] {
    M31_LOCKED
    need(n > 0 && n <= ((size_t)1 << 30), ERR_NOT_POW2, "bad length");
    io.eng.level_for(2 * n);
    m31::F* d = io.in(domain, n);
    m31::F* o = io.eng.tmp(2 * n);
    io.eng.vanish(d, o, n);
    io.out(out, o, 2 * n);
  });
}
// device-pointer variants (inputs and outputs resident in HBM, work enqueued on `stream`)
int ecfft_m31_enter_dev(const ecfft_m31_tree* t, const void* coeffs, size_t n, void* evals, void* stream) {
  return guard31([&] {
    need(t && coeffs && evals, ERR_INVALID_ARG, "null argument");
    DeviceGuard g(t->device);
    m31::set_call_size(n);
    m31::Eng eng(*t, (cudaStream_t)stream);
    eng.enter((const m31::F*)coeffs, (m31::F*)evals, n);
  });
}
int ecfft_m31_exit_dev(const ecfft_m31_tree* t, const void* evals, size_t n, void* coeffs, void* stream) {
  return guard31([&] {
    need(t && coeffs && evals, ERR_INVALID_ARG, "null argument");
    DeviceGuard g(t->device);
    m31::set_call_size(n);
    m31::Eng eng(*t, (cudaStream_t)stream);
    eng.exit((const m31::F*)evals, (m31::F*)coeffs, n);
  });
}
int ecfft_m31_extend_dev(const ecfft_m31_tree* t, const void* evals, size_t n, int moiety, void* out, void* stream) {
  return guard31([&] {
    need(t && evals && out, ERR_INVALID_ARG, "null argument");
    need(moiety == 0 || moiety == 1, ERR_INVALID_ARG, "bad moiety");
    need(n > 0 && n <= ((size_t)1 << 30), ERR_NOT_POW2, "bad length");
    DeviceGuard g(t->device);
    m31::set_call_size(n);
    m31::Eng eng(*t, (cudaStream_t)stream);
    eng.level_for(2 * n);
    eng.extend((const m31::F*)evals, (m31::F*)out, m31::ilog2(n), 1, moiety);
  });
}

}  // extern "C"
