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
static thread_local cudaEvent_t t_open_end = nullptr;   // end event of the launch this host thread has bracketed and not yet closed
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
  t_open_end = r.e1;
  std::lock_guard<std::mutex> lock(g_mu);
  g_recs.push_back(r);
}
void record_end(cudaStream_t st) {   // pairs with this thread's own record_begin, whatever other threads recorded meanwhile
  if (!t_open_end) return;
  ECFFT_CUDA(cudaEventRecord(t_open_end, st));
  t_open_end = nullptr;
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

static inline unsigned grid_for(size_t n, int threads) {
  size_t g = (n + threads - 1) / threads;
  size_t cap = 148u * 64u;  // grid-stride beyond this
  if (g > cap) g = cap;
  if (g == 0) g = 1;
  return (unsigned)g;
}

template <class F>
__global__ void __launch_bounds__(256) k_map(size_t n, F f) {
  // programmatic dependent launch, as in k_extend_sym: the next kernel of the stream may be scheduled while this grid
  // drains, and this one waits here for its own predecessor's results (no-ops without the launch attribute)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) f(i);
}
// ECFFT_B200_MAP_PDL=1: programmatic dependent launch for the glue kernels too.  Off by default: measured mixed
// (profiles/r02_ac_glue_pdl.txt, together with k_extend_sym's: ENTER -> EXIT at 2^12 1.95 -> 1.80 ms, DEGREE 2^16 1.32 -> 1.17 ms,
// but VANISH 2^16 0.84 -> 1.14 ms).
static bool map_pdl() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ECFFT_B200_MAP_PDL");
    v = e ? (atoi(e) != 0) : 0;
  }
  return v != 0;
}
template <class F>
static void map(size_t n, cudaStream_t st, F f) {
  if (n == 0) return;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid_for(n, 256));
  cfg.blockDim = dim3(256);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = map_pdl() ? 1 : 0;
  ECFFT_CUDA(cudaLaunchKernelEx(&cfg, k_map<F>, n, f));
  prof::count_launch();
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

// the same pass for any table set of the unscaled form (the folded combines of Engine::enter_range_serial)
__global__ void __launch_bounds__(256) k_enter_combine_tabs(const Fp* __restrict__ A, const Fp* __restrict__ W, const Fp* __restrict__ xnn,
                                                            const Fp* __restrict__ e0, const Fp* __restrict__ e1, const Fp* __restrict__ o0,
                                                            const Fp* __restrict__ o1, Fp* __restrict__ out, uint32_t log_h, unsigned long long npairs) {
  for (unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; idx < npairs;
       idx += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long blk = idx >> log_h, i = idx & ((1ull << log_h) - 1), off = blk << (log_h + 1), h = 1ull << log_h;
    const Fp u0 = fp_load(A + off + i), v0 = fp_load(A + off + h + i);
    const Fp r0 = e0 ? fp_dot2_lazy(fp_load_ro(e0 + i), u0, fp_load_ro(e1 + i), v0) : fp_muladd_lazy(u0, v0, fp_load_ro(xnn + 2 * i));
    fp_store(out + off + 2 * i, fp_canon(r0));
    const Fp u1 = fp_load(W + off + i), v1 = fp_load(W + off + h + i);
    fp_store(out + off + 2 * i + 1, fp_canon(fp_dot2_lazy(fp_load_ro(o0 + i), u1, fp_load_ro(o1 + i), v1)));
  }
}
void enter_combine_tabs(const SymCombine& c, const Fp* W, uint32_t log_h, size_t n, cudaStream_t st) {
  const size_t npairs = n / 2;
  unsigned grid = (unsigned)((npairs + 255) / 256);
  if (grid > 148u * 32u) grid = 148u * 32u;
  const bool timed = prof::enabled();
  if (timed) prof::record_begin(prof::ENTER_COMBINE, 128.0 * (double)n, st);
  k_enter_combine_tabs<<<grid, 256, 0, st>>>(c.A, W, c.xnn, c.e0, c.e1, c.gam, c.gx, c.out, log_h, npairs);
  if (timed) prof::record_end(st);
  prof::count_launch();
  ECFFT_CUDA(cudaGetLastError());
}
void fold_tables(int group, Fp* t0, Fp* t1, const Fp* gam0, const Fp* gam1, const Fp* gx, const Fp* xnn, const Fp* Pn, Fp two_pow_L, size_t h, cudaStream_t st) {
  switch (group) {
    case 0:  // odd outputs, folded out
      map(h, st, [=] __device__(size_t i) {
        const Fp s = fp_load_ro(Pn + 2 * i + 1);
        fp_store(t0 + i, fp_mul(fp_load_ro(gam1 + i), s));
        fp_store(t1 + i, fp_mul(fp_load_ro(gx + i), s));
      });
      break;
    case 1:  // even outputs, folded in and out: 1 / P[i] = gam0[i] 2^L
      map(h, st, [=] __device__(size_t i) {
        const Fp s = fp_mul(fp_mul_lazy(fp_load_ro(gam0 + i), two_pow_L), fp_load_ro(Pn + 2 * i));
        fp_store(t0 + i, s);
        fp_store(t1 + i, fp_mul(fp_load_ro(xnn + 2 * i), s));
      });
      break;
    case 2:  // folded in, plain out
      map(h, st, [=] __device__(size_t i) {
        const Fp s = fp_mul(fp_load_ro(gam0 + i), two_pow_L);
        fp_store(t0 + i, s);
        fp_store(t1 + i, fp_mul(fp_load_ro(xnn + 2 * i), s));
      });
      break;
    default:  // plain in, folded out
      map(h, st, [=] __device__(size_t i) {
        const Fp s = fp_load_ro(Pn + 2 * i);
        fp_store(t0 + i, s);
        fp_store(t1 + i, fp_mul(fp_load_ro(xnn + 2 * i), s));
      });
  }
}

// ------------------------------------------------------------------------------------------
// pointwise glue
// ------------------------------------------------------------------------------------------
void mul_const(Fp* out, const Fp* in, Fp c, size_t n, cudaStream_t st) {
  map(n, st, [=] __device__(size_t i) { fp_store(out + i, fp_mul(fp_load(in + i), c)); });
}
// out[i] = a[i] * b[i] in the Montgomery domain the API speaks (a~ b~ R^-1 = (ab)~): what ark-ff's `*` on two
// Fp values computes; the caller-side pointwise step of polynomial multiplication (ENTER, multiply, EXIT)
void mul_mont(Fp* out, const Fp* a, const Fp* b, size_t n, cudaStream_t st) {
  const Fp rinv = fp_const_RINV();
  map(n, st, [=] __device__(size_t i) { fp_store(out + i, fp_mul(fp_mul_lazy(fp_load(a + i), fp_load(b + i)), rinv)); });
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
// Tables of the fused REDC (engine.cu): P1 = a0inv * gami_src (* c_even), Kp = -(gam_tgt * a_odd * zinv),
// Zc = zinv (* c_odd); c (MOD's multiplier, fftree.rs:279) may be null.
void redc_tables(Fp* P1, Fp* Kp, Fp* Zc, const Fp* a, const Fp* a0inv, const Fp* zinv, const Fp* gami_src, const Fp* gam_tgt,
                 const Fp* c, size_t h, cudaStream_t st) {
  map(h, st, [=] __device__(size_t i) {
    Fp p1 = fp_mul(fp_load(a0inv + i), fp_load_ro(gami_src + i));
    Fp zc = fp_load(zinv + i);
    Fp kp = fp_mul(fp_mul(fp_load_ro(gam_tgt + i), fp_load(a + 2 * i + 1)), zc);
    if (c) {
      p1 = fp_mul(p1, fp_load(c + 2 * i));
      zc = fp_mul(zc, fp_load(c + 2 * i + 1));
    }
    fp_store(P1 + i, p1);
    fp_store(Kp + i, fp_neg(kp));
    fp_store(Zc + i, zc);
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
// out[w][2i] = Q[2w][i] * Q[2w+1][i]: the sibling product written where VANISH's interleave wants it (src/fftree.rs:303-307)
void mul_pairs_even(Fp* out, const Fp* Q, size_t len, size_t npairs, cudaStream_t st) {
  map(len * npairs, st, [=] __device__(size_t idx) {
    size_t w = idx / len, i = idx % len;
    fp_store(out + w * 2 * len + 2 * i, fp_mul(fp_load(Q + 2 * w * len + i), fp_load(Q + (2 * w + 1) * len + i)));
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
void degree_step(const unsigned long long* diff, Fp* e1, const Fp* g1, const Fp* zinv, const Fp* e0, Fp* next, size_t h,
                 unsigned long long* result, cudaStream_t st) {
  map(h, st, [=] __device__(size_t i) {
    if (__ldcg(diff) == 0) {
      fp_store(next + i, fp_load(e0 + i));
    } else {
      fp_store(e1 + i, fp_mul(fp_sub(fp_load(e1 + i), fp_load(g1 + i)), fp_load_ro(zinv + i)));
      if (i == 0) atomicAdd(result, (unsigned long long)h);
    }
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
// A zero denominator (the reference's `rational_map.map(..).unwrap()` panics, src/fftree.rs:57-58) is counted in *err.
void ratmap_layer(Fp* layer, const Fp* prev, size_t count, const Fp* num, int nnum, const Fp* den, int nden, unsigned long long* err, cudaStream_t st) {
  map(count, st, [=] __device__(size_t i) {
    Fp x = fp_load(prev + i);
    Fp nu = poly_eval_dev(num, nnum, x), de = poly_eval_dev(den, nden, x);
    if (err && fp_is_zero(de)) atomicAdd(err, 1ull);
    fp_store(layer + i, fp_mul(nu, fp_inv(de)));
  });
}
// Lemma 3.2 matrices of one layer, reference src/fftree.rs:354-362.  flayer has 2d entries at
// stride fstride (the chain level's f layer is a strided view of the top tree's).
// A singular matrix (the reference's `rmat.inverse().unwrap()` panics, src/fftree.rs:361) is counted in *err.
void build_matrices(Fp* rl, Fp* dl, const Fp* flayer, size_t fstride, size_t d, const Fp* den, int nden, unsigned long long* err, cudaStream_t st) {
  uint64_t e = d / 2 - 1;
  map(d, st, [=] __device__(size_t i) {
    Fp s0 = fp_load(flayer + i * fstride), s1 = fp_load(flayer + (i + d) * fstride);
    Fp v0 = fp_pow_u64(poly_eval_dev(den, nden, s0), e);
    Fp v1 = fp_pow_u64(poly_eval_dev(den, nden, s1), e);
    Fp r0 = v0, r1 = fp_mul(s0, v0), r2 = v1, r3 = fp_mul(s1, v1);
    Fp det = fp_sub(fp_mul(r0, r3), fp_mul(r1, r2));
    if (err && fp_is_zero(det)) atomicAdd(err, 1ull);
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

// ------------------------------------------------------------------------------------------
// multi-GPU building blocks: one butterfly level whose pairs straddle two ranks, and the
// ENTER combine on a slice.  `own`/`partner` are the two ranks' chunks of the same vector, element e of
// both belonging to the same pair; role 0: own holds the lower (p) element, 1: the upper (q) element.
// ------------------------------------------------------------------------------------------
void mg_cross(const Level& lv, int phase, uint32_t j, int role, size_t p_pos0, const Fp* own, const Fp* partner, size_t count, Fp* out, cudaStream_t st,
              Moiety source, Moiety target) {
  const Fp* table = phase == 0 ? lv.tw_d[source] : lv.tw_r[target];
  if (!table) throw Error(ERR_MISSING_TABLES, "mg_cross: normalised tables missing");
  const size_t mask = ((size_t)1 << j) - 1, ibase = p_pos0 & mask;
  if (lv.sym) {  // symmetric form: one twiddle per pair
    const Fp* layer = table + ((size_t)1 << j);
    map(count, st, [=] __device__(size_t e) {
      Fp g = fp_load_ro(layer + ((ibase + e) & mask));
      Fp xo = fp_load(own + e), xr = fp_load(partner + e);
      Fp xp = role == 0 ? xo : xr, xq = role == 0 ? xr : xo;
      Fp res;
      if (phase == 0)  // decompose: x_p = y_p + y_q, x_q = (y_p - y_q)/g
        res = role == 0 ? fp_add_lazy(xp, xq) : fp_mul_lazy(g, fp_sub_lazy2(xp, xq));
      else {           // recombine: y_p = x_p + g x_q, y_q = x_p - g x_q
        Fp t = fp_mul_lazy(g, xq);
        res = role == 0 ? fp_add_lazy(xp, t) : fp_sub_lazy2(xp, t);
      }
      fp_store(out + e, fp_canon(res));
    });
    return;
  }
  const Fp* layer = table + 2 * ((size_t)1 << j);
  map(count, st, [=] __device__(size_t e) {
    const Fp* tw = layer + 2 * ((ibase + e) & mask);
    Fp xo = fp_load(own + e), xr = fp_load(partner + e);
    Fp xp = role == 0 ? xo : xr, xq = role == 0 ? xr : xo;
    Fp res;
    if (phase == 0)    // decompose, sum form: y_q = x^_p + x^_q, y_p = -(s1 x^_p + s0 x^_q)
      res = role == 1 ? fp_add_lazy(xp, xq) : fp_dot2_lazy(fp_load_ro(tw), xp, fp_load_ro(tw + 1), xq);
    else               // recombine: y_p = x_p + s0 x_q, y_q = x_p + s1 x_q
      res = fp_muladd_lazy(xp, fp_load_ro(tw + role), xq);
    fp_store(out + e, fp_canon(res));
  });
}
// Stream-ordered flags in peer-mapped arenas (include/ecfft_b200.h "peer exchange").  One thread: publish
// `value` in own_flag (release at system scope: everything enqueued before is visible to the node), then
// spin until each given peer flag shows >= value.  A wait not satisfied within the timeout traps — a CUDA
// error on the next call instead of a hung GPU.
// A wait that is not satisfied within the timeout records {1<<63 | value << 24 | info << 8 | which} in *status (the
// waiting rank's own arena; first record wins; ecfft_mg_arena_status reads it) and then traps — unless
// ECFFT_B200_PEER_NO_TRAP is set, in which case the kernel gives up waiting and the call's results are garbage
// that the caller detects through the status word.
__device__ __forceinline__ void mg_timeout(unsigned long long* status, unsigned long long value, unsigned info, unsigned which, int no_trap) {
  if (status) {
    atomicCAS(status, 0ull, (1ull << 63) | (value << 24) | ((unsigned long long)(info & 0xffff) << 8) | (which & 0xff));
    __threadfence_system();
  }
  if (!no_trap) __trap();
}
__global__ void k_mg_sync(unsigned long long* own_flag, unsigned long long value, const unsigned long long* wait_a,
                          const unsigned long long* wait_b, unsigned long long timeout_ns, unsigned long long* status, unsigned info, int no_trap) {
  if (own_flag) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(own_flag), "l"(value) : "memory");
  }
  const unsigned long long* w[2] = {wait_a, wait_b};
  unsigned long long t0, t, v;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (int i = 0; i < 2; i++) {
    if (!w[i]) continue;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(w[i]) : "memory");
      if (v >= value) break;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (timeout_ns && t - t0 > timeout_ns) {  // 0 = wait for ever
        mg_timeout(status, value, info, (unsigned)i, no_trap);
        break;
      }
      __nanosleep(100);
    }
  }
  __threadfence_system();
}
// all-peers form: spin until flag[idx] of every other rank's arena shows >= value
struct ArenaBases { const unsigned long long* base[16]; };
__global__ void k_mg_wait_all(ArenaBases b, int world, int rank, unsigned idx, unsigned long long value, unsigned long long timeout_ns,
                              unsigned long long* status, int no_trap) {
  unsigned long long t0, t, v;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (int r = 0; r < world; r++) {
    if (r == rank) continue;
    for (;;) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(b.base[r] + idx) : "memory");
      if (v >= value) break;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      if (timeout_ns && t - t0 > timeout_ns) {  // 0 = wait for ever
        mg_timeout(status, value, 0xffffu, (unsigned)r, no_trap);
        break;
      }
      __nanosleep(100);
    }
  }
  __threadfence_system();
}
static int peer_no_trap() {
  static int v = -1;
  if (v < 0) v = getenv("ECFFT_B200_PEER_NO_TRAP") != nullptr;
  return v;
}
void mg_wait_all(void* const* bases, int world, int rank, unsigned idx, unsigned long long value, unsigned timeout_ms, cudaStream_t st,
                 unsigned long long* status) {
  if (world > 16) throw Error(ERR_INVALID_ARG, "peer schedule supports at most 16 ranks");
  ArenaBases b{};
  for (int r = 0; r < world; r++) b.base[r] = (const unsigned long long*)bases[r];
  k_mg_wait_all<<<1, 1, 0, st>>>(b, world, rank, idx, value, (unsigned long long)timeout_ms * 1000000ull, status, peer_no_trap());
  prof::count_launch();
  ECFFT_CUDA(cudaGetLastError());
}
void mg_sync(unsigned long long* own_flag, unsigned long long value, const unsigned long long* wait_a,
             const unsigned long long* wait_b, unsigned timeout_ms, cudaStream_t st, unsigned long long* status, unsigned info) {
  k_mg_sync<<<1, 1, 0, st>>>(own_flag, value, wait_a, wait_b, (unsigned long long)timeout_ms * 1000000ull, status, info, peer_no_trap());
  prof::count_launch();
  ECFFT_CUDA(cudaGetLastError());
}

// out[2t] = u0[t] + v0[t]*xnn[2(i0+t)], out[2t+1] = gam1[i0+t]*u1[t] + gx[i0+t]*v1[t]  (u1, v1 unscaled)
void mg_combine(const Level& lv, size_t i0, const Fp* u0, const Fp* v0, const Fp* u1, const Fp* v1, size_t count, Fp* out, cudaStream_t st) {
  if (!lv.gx || !lv.gam[1]) throw Error(ERR_MISSING_TABLES, "mg_combine: normalised tables missing");
  const Fp* xnn = lv.xnn_s + 2 * i0;
  const Fp* gam = lv.gam[1] + i0;
  const Fp* gx = lv.gx + i0;
  map(count, st, [=] __device__(size_t t) {
    fp_store(out + 2 * t, fp_canon(fp_muladd_lazy(fp_load(u0 + t), fp_load(v0 + t), fp_load_ro(xnn + 2 * t))));
    fp_store(out + 2 * t + 1, fp_canon(fp_dot2_lazy(fp_load_ro(gam + t), fp_load(u1 + t), fp_load_ro(gx + t), fp_load(v1 + t))));
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
    fp_store(tw_d + 2 * idx, fp_neg(s1));
    fp_store(tw_d + 2 * idx + 1, fp_neg(s0));
  });
}
// Sum-form decompose: fold the per-level input scalings (-c for the lower, +c for the upper element of
// each pair, c = 1/(s1-s0) of that level's pair) into the pre-scale table: gami[p] *= prod_j (+-c_j(p)).
void fold_sumform_prescale(Fp* gami, const Fp* f_top, size_t fstride, size_t h, int mu, cudaStream_t st) {
  map(h, st, [=] __device__(size_t p) {
    Fp acc = fp_one();
    bool neg = false;
    for (uint32_t j = 0; ((size_t)1 << j) < h; j++) {
      size_t i = p & (((size_t)1 << j) - 1), B = (size_t)2 << j;
      Fp s0 = fp_load(f_top + (2 * B + 2 * i + mu) * fstride);
      Fp s1 = fp_load(f_top + (2 * B + 2 * i + mu + B) * fstride);
      acc = fp_mul(acc, fp_sub(s1, s0));
      if (((p >> j) & 1) == 0) neg = !neg;
    }
    Fp c = fp_inv(acc);  // product of the c_j
    if (neg) c = fp_neg(c);
    fp_store(gami + p, fp_mul(fp_load(gami + p), c));
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
// Symmetric-form tables (DESIGN.md 4.1 "symmetric"): the level's map x -> (x^2 + c1 x + beta^2)/x
// identifies s with beta^2/s, and g(s) = (s - beta)/(s + beta) takes opposite values on the two.
// Entry idx = 2^j + i holds g(s0) resp. 1/g(s0) for s0 = the pair's lower node.
void build_twiddles_sym(Fp* tw_r, Fp* tw_d, const Fp* f_top, size_t fstride, size_t h, int mu, const Fp* beta_by_j, unsigned long long* err, cudaStream_t st) {
  map(h, st, [=] __device__(size_t idx) {
    if (idx == 0) {
      fp_store(tw_r, fp_zero());
      fp_store(tw_d, fp_zero());
      return;
    }
    uint32_t j = 63 - __clzll((unsigned long long)idx);
    size_t i = idx - ((size_t)1 << j), B = (size_t)2 << j;
    Fp s0 = fp_load(f_top + (2 * B + 2 * i + mu) * fstride);
    Fp b = fp_load_ro(beta_by_j + j);
    Fp nu = fp_sub(s0, b), de = fp_add(s0, b);
    if (err && (fp_is_zero(nu) || fp_is_zero(de))) atomicAdd(err, 1ull);  // a node at a fixed point +-beta of the involution
    Fp t = fp_inv(fp_mul(nu, de));
    fp_store(tw_r + idx, fp_mul(fp_mul(nu, nu), t));
    fp_store(tw_d + idx, fp_mul(fp_mul(de, de), t));
  });
}
// Gamma^mu_p = prod_j (s + beta_j) v(s)^(2^j - 1) over the nodes s the position passes through; the v
// powers are the first-column entries of the recombine matrices as in build_gamma.
void build_gamma_sym(Fp* gam, const Fp* rmat, const Fp* f_top, size_t fstride, size_t h, int mu, const Fp* beta_by_j, cudaStream_t st) {
  map(h, st, [=] __device__(size_t p) {
    Fp acc = fp_one();
    for (uint32_t j = 0; ((size_t)1 << j) < h; j++) {
      size_t i = p & (((size_t)1 << j) - 1), b = (p >> j) & 1, B = (size_t)2 << j;
      Fp s = fp_load(f_top + (2 * B + 2 * i + mu + b * B) * fstride);
      acc = fp_mul_lazy(acc, fp_load(rmat + 4 * (B + 2 * i + mu) + 2 * b));
      acc = fp_mul_lazy(acc, fp_add(s, fp_load_ro(beta_by_j + j)));
    }
    fp_store(gam + p, fp_canon(acc));
  });
}
// ------------------------------------------------------------------------------------------
// Device self-test of the lazy add / subtract forms the butterflies use (fp_add_lazy_f, fp_sub_lazy2_f and
// their branch-free siblings): directed operands that drive the carry out of the low two limbs ("ripple")
// and the second wrap, compared with canonical arithmetic on the canonicalised operands.  The host build of
// fp.cuh is fuzzed the same way (tests/test_fp_host.py); this runs the inline-PTX carry chains themselves.
// counters: [0] mismatches, [1] additions that took the ripple path, [2] subtractions that did.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long selftest_mix(unsigned long long& x) {
  x += 0x9E3779B97F4A7C15ull;
  unsigned long long z = x;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ Fp selftest_wrap_sub(const Fp& a, const Fp& b) {  // a - b mod 2^256
  Fp d;
  d.v[0] = sub_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) d.v[i] = subc_cc(a.v[i], b.v[i]);
  return d;
}
__global__ void __launch_bounds__(256) k_selftest_addsub(unsigned long long* counters, unsigned long long n) {
  for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (unsigned long long)gridDim.x * blockDim.x) {
    unsigned long long st = t * 0x2545F4914F6CDD1Dull + 12345;
    Fp a, tgt;
    for (int i = 0; i < 4; i++) {
      unsigned long long r = selftest_mix(st), q = selftest_mix(st);
      a.v[2 * i] = (uint32_t)r; a.v[2 * i + 1] = (uint32_t)(r >> 32);
      tgt.v[2 * i] = (uint32_t)q; tgt.v[2 * i + 1] = (uint32_t)(q >> 32);
    }
    const unsigned kind = (unsigned)(t % 6);
    const unsigned long long small = selftest_mix(st) >> (31 + (selftest_mix(st) & 31));  // 1..33 significant bits
    if (kind == 0 || kind == 1) {        // wrapped sum with low 64 bits within DELTA of 2^64 (kind 1: high limbs all ones)
      const unsigned long long lo = ~0ull - small;
      tgt.v[0] = (uint32_t)lo; tgt.v[1] = (uint32_t)(lo >> 32);
      if (kind == 1) for (int i = 2; i < 8; i++) tgt.v[i] = 0xFFFFFFFFu;
    } else if (kind == 2 || kind == 3) { // borrowed difference with low 64 bits below DELTA (kind 3: high limbs zero)
      tgt.v[0] = (uint32_t)small; tgt.v[1] = (uint32_t)(small >> 32);
      if (kind == 3) for (int i = 2; i < 8; i++) tgt.v[i] = 0u;
    } else if (kind == 4) {              // operands near 2^256
      for (int i = 2; i < 8; i++) a.v[i] = 0xFFFFFFFFu;
    }
    // kinds 0,1: b = tgt - a (so a + b = tgt mod 2^256); kinds 2,3: b = a - tgt (so a - b = tgt mod 2^256)
    const Fp b = (kind <= 1) ? selftest_wrap_sub(tgt, a) : (kind <= 3 ? selftest_wrap_sub(a, tgt) : tgt);
    const Fp ac = fp_canon(a), bc = fp_canon(b);
    const Fp want_add = fp_add(ac, bc), want_sub = fp_sub(ac, bc);
    unsigned long long bad = 0;
    bad += !fp_eq(fp_canon(fp_add_lazy_f(a, b)), want_add);
    bad += !fp_eq(fp_canon(fp_add_lazy(a, b)), want_add);
    bad += !fp_eq(fp_canon(fp_sub_lazy2_f(a, b)), want_sub);
    bad += !fp_eq(fp_canon(fp_sub_lazy2(a, b)), want_sub);
    if (bad) atomicAdd(counters, bad);
    // did this sample take the ripple paths?  (carry/borrow out of the low two limbs after the DELTA fold)
    {
      Fp s;
      s.v[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
      for (int i = 1; i < 8; i++) s.v[i] = addc_cc(a.v[i], b.v[i]);
      const uint32_t c = addc(0u, 0u);
      const unsigned long long lo = ((unsigned long long)s.v[1] << 32) | s.v[0];
      if (c && lo + 0x1000003D1ull < lo) atomicAdd(counters + 1, 1ull);
      const Fp d = selftest_wrap_sub(a, b);
      bool borrow = false;  // a < b as 256-bit integers
      for (int i = 7; i >= 0; i--) {
        if (a.v[i] != b.v[i]) { borrow = a.v[i] < b.v[i]; break; }
      }
      const unsigned long long dlo = ((unsigned long long)d.v[1] << 32) | d.v[0];
      if (borrow && dlo < 0x1000003D1ull) atomicAdd(counters + 2, 1ull);
    }
  }
}
void selftest_field(unsigned long long* counters, unsigned long long n, cudaStream_t st) {
  k_selftest_addsub<<<148 * 8, 256, 0, st>>>(counters, n);
  prof::count_launch();
  ECFFT_CUDA(cudaGetLastError());
}
// out[i] = e[i * e_stride] * z[i] + g[i] * kp[i]   (REDC's h1 on a chunk, fftree.rs:253-255 with the folded tables)
void dot2_strided(Fp* out, const Fp* e, size_t e_stride, const Fp* z, const Fp* g, const Fp* kp, size_t n, cudaStream_t st) {
  map(n, st, [=] __device__(size_t i) {
    fp_store(out + i, fp_canon(fp_dot2_lazy(fp_load(e + i * e_stride), fp_load_ro(z + i), fp_load(g + i), fp_load_ro(kp + i))));
  });
}
// out[i] = (a[i * a_stride] - b[i]) * c[i]   (EXIT's v0 on a chunk, fftree.rs:215-219)
void sub_mul_strided(Fp* out, const Fp* a, size_t a_stride, const Fp* b, const Fp* c, size_t n, cudaStream_t st) {
  map(n, st, [=] __device__(size_t i) { fp_store(out + i, fp_mul(fp_sub(fp_load(a + i * a_stride), fp_load(b + i)), fp_load_ro(c + i))); });
}
void mul_strided(Fp* out, const Fp* a, const Fp* b, size_t b_stride, size_t b_off, size_t n, cudaStream_t st) {
  map(n, st, [=] __device__(size_t i) { fp_store(out + i, fp_mul(fp_load(a + i), fp_load(b + b_off + i * b_stride))); });
}

}  // namespace k
}  // namespace ecfft
