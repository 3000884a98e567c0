// The FFTree<F> algorithms (reference src/fftree.rs:72-316) as host-side schedules of batched
// device kernels.  The reference recurses on Vec<F>; here every recursion depth is one batched
// launch over all sub-problems of that depth (they share the chain level's tables):
//   ENTER / VANISH run bottom-up, EXIT runs top-down, DEGREE follows its single branch.
#include <cstdlib>
#include <utility>

#include "engine.h"

namespace ecfft {

static inline bool is_pow2(size_t n) { return n && !(n & (n - 1)); }
static inline uint32_t ilog2(size_t n) {
  uint32_t l = 0;
  while (n >>= 1) l++;
  return l;
}

Tree::~Tree() {
  int prev = -1;
  cudaGetDevice(&prev);
  cudaSetDevice(device);
  if (build_errors) cudaFree(build_errors);
  for (void* p : owned) cudaFree(p);
  for (void* p : owned_lazy) cudaFree(p);
  if (stream) cudaStreamDestroy(stream);
  for (cudaStream_t s : aux) cudaStreamDestroy(s);
  if (prev >= 0 && prev != device) cudaSetDevice(prev);
}
Fp* Tree::dalloc(size_t count) {
  void* p = nullptr;
  ECFFT_CUDA(cudaMalloc(&p, (count ? count : 1) * sizeof(Fp)));
  owned.push_back(p);
  return (Fp*)p;
}

// subtree_with_size, reference src/fftree.rs:489-496
const Level& Engine::level_for(size_t leaves) const {
  if (!is_pow2(leaves)) throw Error(ERR_NOT_POW2, "length is not a power of two");
  uint32_t lg = ilog2(leaves);
  if (lg > t.log_n) throw Error(ERR_TREE_TOO_SMALL, "FFTree is too small");
  return t.levels[lg];
}

Fp* Engine::tmp(size_t count) const {
  void* p = nullptr;
  ECFFT_CUDA(cudaMallocAsync(&p, (count ? count : 1) * sizeof(Fp), st));
  return (Fp*)p;
}
void Engine::release(Fp* p) const {
  if (p) ECFFT_CUDA(cudaFreeAsync(p, st));
}

// FFTree::extend, src/fftree.rs:123-126 (batched)
void Engine::extend(const Fp* in, Fp* out, size_t h, size_t nvec, Moiety target) const {
  const Level& lv = level_for(h * 2);
  k::extend(lv, in, out, ilog2(h), nvec, target, st);
}

// FFTree::enter_impl, src/fftree.rs:143-161, flattened bottom-up: after the pass for m the
// array holds n/m evaluation vectors of length m (one per coefficient chunk).
// Number of concurrent streams ENTER spreads independent coefficient ranges over (ECFFT_B200_ENTER_STREAMS,
// default 2): the ranges do not interact below block size n/S, and kernels of different streams fill each
// other's end-of-launch drain (a launch loses about half a CTA lifetime of SM occupancy while it drains).
// Default: two streams from 2^21 elements up, four for 2^19 .. 2^20 (launches of such sizes do not fill the GPU
// on their own: measured 1.97 -> 1.89 ms at 2^19, 3.48 -> 3.42 ms at 2^20, but 14.7 -> 14.9 ms at 2^22;
// profiles/r02_e_*, r02_f_*).
static int enter_streams(size_t n = (size_t)1 << 22) {
  static int s = -1;
  if (s < 0) {
    const char* e = getenv("ECFFT_B200_ENTER_STREAMS");
    s = e ? atoi(e) : 0;
    if (s != 1 && s != 2 && s != 4) s = 0;
  }
  return s ? s : (n <= ((size_t)1 << 20) ? 4 : 2);
}
// smallest range (elements per stream) worth a stream of its own (ECFFT_B200_ENTER_FORK_MIN = log2, default 17)
static size_t enter_fork_min() {
  static size_t v = 0;
  if (!v) {
    const char* e = getenv("ECFFT_B200_ENTER_FORK_MIN");
    int lg = e ? atoi(e) : 17;
    if (lg < 10 || lg > 40) lg = 17;
    v = (size_t)1 << lg;
  }
  return v;
}
namespace {
struct ScopedEvent {  // cudaEventDestroy defers the release until the recorded work has completed
  cudaEvent_t e = nullptr;
  ScopedEvent() { ECFFT_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); }
  ~ScopedEvent() { if (e) cudaEventDestroy(e); }
  ScopedEvent(const ScopedEvent&) = delete;
  ScopedEvent& operator=(const ScopedEvent&) = delete;
};
}  // namespace
cudaStream_t Tree::aux_stream(int i) const {
  std::lock_guard<std::mutex> lock(aux_mu);
  while ((int)aux.size() <= i) {
    cudaStream_t s = nullptr;
    ECFFT_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    aux.push_back(s);
  }
  return aux[i];
}

// ECFFT_B200_FOLD (default 1): between two depths of one ENTER the data is stored already multiplied by the
// pre-scale of the EXTEND that reads it next.  That EXTEND then skips its pre-scale pass (one product and one
// reduction per element), the combine that produced the data pays nothing for it (its even output becomes a
// two-product dot with the scale in the tables, its odd output already was one), and the combine after it
// divides the scale back out through its own tables: 2.5 -> 2 products and 2 -> 1 reductions per element and
// depth outside the butterflies.  The first depth of a range reads plain data, the last one writes plain data.
static bool fold_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ECFFT_B200_FOLD");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v != 0;
}
void Engine::enter_range(const Fp* in, Fp* out, size_t n, size_t m_lo, size_t m_hi, int streams_hint) const {
  // n is any whole number of m_hi-blocks (the blocks are independent): a power of two for a full ENTER
  if (!is_pow2(m_lo) || !is_pow2(m_hi)) throw Error(ERR_NOT_POW2, "length is not a power of two");
  if (n == 0 || m_hi > n || m_lo > m_hi || n % m_hi) throw Error(ERR_INVALID_ARG, "enter: bad level range");
  level_for(m_hi);
  if (m_lo == m_hi) {
    if (in != out) ECFFT_CUDA(cudaMemcpyAsync(out, in, n * sizeof(Fp), cudaMemcpyDeviceToDevice, st));
    return;
  }
  if (enter_range_flow(in, out, n, m_lo, m_hi)) return;
  // Independent ranges on concurrent streams: range s runs the depths up to m_mid (the largest block size
  // that tiles a range) on stream s, the caller's stream joins them and runs the remaining depths.
  // streams_hint: the caller's measured choice where the automatic one does not apply (the rank-local ENTER of the
  // sharded schedule, whose launches are followed by flag waits instead of the next call's work); the environment wins
  const int S = (streams_hint > 0 && getenv("ECFFT_B200_ENTER_STREAMS") == nullptr) ? streams_hint : enter_streams(n);
  if (S > 1 && !prof::enabled() && n % (size_t)S == 0 && n / (size_t)S >= enter_fork_min()) {
    const size_t part = n / (size_t)S;
    size_t m_mid = m_hi;
    while (m_mid > m_lo && part % m_mid) m_mid /= 2;
    if (m_mid > m_lo) {
      Fp* mid = m_mid == m_hi ? out : tmp(n);
      // the data handed from the per-range depths to the joint depths stays folded (enter_range_serial)
      bool keep = m_mid < m_hi && fold_enabled() && k::butterfly_mode() == 2 && getenv("ECFFT_B200_NO_COMBINE_FUSION") == nullptr;
      for (size_t m = m_lo * 2; keep && m <= m_hi; m *= 2) keep = level_for(m).sym && level_for(m).has_norm();
      ScopedEvent fork, join[3];
      ECFFT_CUDA(cudaEventRecord(fork.e, st));
      for (int s = 1; s < S; s++) {
        cudaStream_t as = t.aux_stream(s - 1);
        ECFFT_CUDA(cudaStreamWaitEvent(as, fork.e, 0));
        Engine sub(t, as);
        sub.enter_range_serial(in + s * part, mid + s * part, part, m_lo, m_mid, false, keep);
        ECFFT_CUDA(cudaEventRecord(join[s - 1].e, as));
      }
      enter_range_serial(in, mid, part, m_lo, m_mid, false, keep);
      for (int s = 1; s < S; s++) ECFFT_CUDA(cudaStreamWaitEvent(st, join[s - 1].e, 0));
      if (m_mid < m_hi) {
        enter_range_serial(mid, out, n, m_mid, m_hi, keep, false);
        release(mid);
      }
      return;
    }
  }
  enter_range_serial(in, out, n, m_lo, m_hi);
}

// All depths m_lo < m <= m_hi as ONE flow launch (sym_kernel.cu: persistent CTAs, per-block dependency counters
// instead of kernel boundaries).  The work buffers alternate by depth parity: everything a depth reads or
// writes lies inside its own aligned m-block, and a tile of depth d+1 only starts once the depth-d results of
// its whole block are complete, so a buffer is never overwritten while a tile still needs its old contents.
bool Engine::enter_range_flow(const Fp* in, Fp* out, size_t n, size_t m_lo, size_t m_hi) const {
  if (!k::flow_enabled() || k::butterfly_mode() != 2) return false;
  const size_t T = (size_t)1 << k::flow_log_tile();
  if (n % T || m_lo == m_hi) return false;
  for (size_t m = m_lo * 2; m <= m_hi; m *= 2) {
    const Level& lv = level_for(m);
    if (!(lv.sym && lv.has_norm())) return false;
  }
  Fp* W[2] = {nullptr, nullptr};
  Fp* ping[2] = {nullptr, nullptr};
  k::SymFlow flow;
  const Fp* cur = in;
  uint32_t idx = 0;
  for (size_t m = m_lo * 2; m <= m_hi; m *= 2, idx++) {
    const Level& lv = level_for(m);
    const size_t h = m / 2;
    const uint32_t log_h = ilog2(h);
    Fp* dst;
    if (m == m_hi && out != in) {
      dst = out;
    } else {
      if (!ping[idx & 1]) ping[idx & 1] = tmp(n);
      dst = ping[idx & 1];
    }
    k::SymCombine c{cur, lv.xnn_s, lv.gam[1], lv.gx, dst};
    if (h == 1) {
      k::plan_combine_only(flow, c, cur, 0, n);   // EXTEND of a length-1 vector is the identity (fftree.rs:74-76)
    } else {
      if (log_h + 1 > k::flow_log_tile() && !W[idx & 1]) W[idx & 1] = tmp(n);
      Fp* Wd = W[idx & 1];
      if (!k::plan_extend_sym(flow, lv.tw_d[0], lv.tw_r[1], lv.ctr[1], cur, Wd, log_h, n / h, lv.gami[0], nullptr, &c)) {
        // a vector fills the tile: EXTEND (unscaled) into the work buffer, then the combine as its own pass
        if (!k::plan_extend_sym(flow, lv.tw_d[0], lv.tw_r[1], lv.ctr[1], cur, Wd, log_h, n / h, lv.gami[0], nullptr, nullptr))
          throw Error(ERR_INVALID_ARG, "enter: EXTEND refused a depth it should take");
        k::plan_combine_only(flow, c, Wd, log_h, n);
      }
    }
    cur = dst;
  }
  k::launch_flow(flow, st);
  if (cur != out) ECFFT_CUDA(cudaMemcpyAsync(out, cur, n * sizeof(Fp), cudaMemcpyDeviceToDevice, st));
  release(W[0]);
  release(W[1]);
  release(ping[0]);
  release(ping[1]);
  return true;
}

bool Engine::fold_tabs(const Level& lv, int group) const {
  const uint32_t k = lv.log_n;
  if (k < 1 || !(lv.sym && lv.has_norm())) return false;
  const bool needs_next = group != 2;
  if (needs_next && (k + 1 > t.log_n || !t.levels[k + 1].gami[0])) return false;
  std::lock_guard<std::mutex> lock(t.tab_mu);
  if (lv.fold_tab[2 * group + 1]) return true;
  const size_t h = (size_t)1 << (k - 1);
  cudaStream_t bs = t.stream;
  Fp* tab[2];
  for (int i = 0; i < 2; i++) {
    void* p = nullptr;
    ECFFT_CUDA(cudaMallocAsync(&p, h * sizeof(Fp), bs));   // stream-ordered, see exit_tabs
    t.owned_lazy.push_back(p);
    tab[i] = (Fp*)p;
  }
  Fp two_pow_L = fp_zero();   // 1 / gami[0][i] = gam[0][i] 2^(k-1) (build_norm_tables)
  two_pow_L.v[(k - 1) / 32] = 1u << ((k - 1) % 32);
  k::fold_tables(group, tab[0], tab[1], lv.gam[0], lv.gam[1], lv.gx, lv.xnn_s, needs_next ? t.levels[k + 1].gami[0] : nullptr, two_pow_L, h, bs);
  ECFFT_CUDA(cudaStreamSynchronize(bs));
  lv.fold_tab[2 * group] = tab[0];
  lv.fold_tab[2 * group + 1] = tab[1];
  return true;
}

void Engine::enter_range_serial(const Fp* in, Fp* out, size_t n, size_t m_lo, size_t m_hi, bool in_folded, bool out_folded) const {
  if (m_lo == m_hi) {
    if (in_folded != out_folded) throw Error(ERR_INVALID_ARG, "enter: an empty range cannot change the folding");
    if (in != out) ECFFT_CUDA(cudaMemcpyAsync(out, in, n * sizeof(Fp), cudaMemcpyDeviceToDevice, st));
    return;
  }
  // folding needs the symmetric tables on every depth of the range (and the level above a folded output)
  bool fold = fold_enabled() && k::butterfly_mode() == 2 && getenv("ECFFT_B200_NO_COMBINE_FUSION") == nullptr;
  for (size_t m = m_lo * 2; fold && m <= m_hi; m *= 2) {
    const Level& lv = level_for(m);
    fold = lv.sym && lv.has_norm();
  }
  if (!fold && (in_folded || out_folded)) throw Error(ERR_INVALID_ARG, "enter: folded data needs the symmetric tables");
  Fp* W = tmp(n);
  Fp* ping[2] = {nullptr, nullptr};
  const Fp* cur = in;
  uint32_t idx = 0;
  for (size_t m = m_lo * 2; m <= m_hi; m *= 2, idx++) {
    const Level& lv = level_for(m);
    const size_t h = m / 2;
    Fp* dst;
    if (m == m_hi && out != in) {
      dst = out;
    } else {
      if (!ping[idx & 1]) ping[idx & 1] = tmp(n);
      dst = ping[idx & 1];
    }
    // u1, v1 for every block at once; with the normalised tables the Gamma^1 scaling of the
    // EXTEND output is folded into the combine
    const bool unscaled = k::butterfly_mode() != 0 && lv.has_norm();
    static const bool no_fuse = getenv("ECFFT_B200_NO_COMBINE_FUSION") != nullptr;
    if (unscaled && lv.sym && !no_fuse) {  // EXTEND with the combine fused into its last pass
      bool fin = fold && (m == m_lo * 2 ? in_folded : true);
      bool fout = fold && (m == m_hi ? out_folded : true);
      if (fout && !fold_tabs(lv, 0)) throw Error(ERR_INVALID_ARG, "enter: no level above to fold for");
      if (fin && fout && !fold_tabs(lv, 1)) throw Error(ERR_INVALID_ARG, "enter: fold tables");
      if (fin && !fout && !fold_tabs(lv, 2)) throw Error(ERR_INVALID_ARG, "enter: fold tables");
      if (!fin && fout && !fold_tabs(lv, 3)) throw Error(ERR_INVALID_ARG, "enter: fold tables");
      k::SymCombine c{cur, lv.xnn_s, fout ? lv.fold_tab[0] : lv.gam[1], fout ? lv.fold_tab[1] : lv.gx, dst};
      if (fin || fout) {
        const int g = fin && fout ? 1 : (fin ? 2 : 3);
        c.e0 = lv.fold_tab[2 * g];
        c.e1 = lv.fold_tab[2 * g + 1];
      }
      const Fp* pre = fin ? nullptr : lv.gami[0];
      const uint32_t log_h = ilog2(h);
      if (!k::extend_sym(lv.tw_d[0], lv.tw_r[1], lv.ctr[1], cur, W, log_h, n / h, pre, nullptr, &c, st)) {
        // depths the fused pass does not take (h = 1: EXTEND is the identity, fftree.rs:74-76; h = tile: no room for
        // the sibling vector): unscaled EXTEND into W, then the combine as a pass of its own
        const Fp* Wsrc = cur;   // h = 1: both scales of the 2-leaf level are 1
        if (h > 1) {
          if (!k::extend_sym(lv.tw_d[0], lv.tw_r[1], lv.ctr[1], cur, W, log_h, n / h, pre, nullptr, nullptr, st)) {
            k::extend(lv, cur, W, log_h, n / h, S1, st, true);   // fewer than 4 elements: the radix-2 kernel (applies the pre-scale itself)
            if (fin) throw Error(ERR_INVALID_ARG, "enter: EXTEND refused a depth it should take");
          }
          Wsrc = W;
        }
        k::enter_combine_tabs(c, Wsrc, log_h, n, st);
      }
      cur = dst;
      continue;
    }
    k::extend(lv, cur, W, ilog2(h), n / h, S1, st, unscaled);
    k::enter_combine(lv, cur, W, dst, ilog2(h), n, unscaled, st);
    cur = dst;
  }
  if (cur != out) ECFFT_CUDA(cudaMemcpyAsync(out, cur, n * sizeof(Fp), cudaMemcpyDeviceToDevice, st));
  release(W);
  release(ping[0]);
  release(ping[1]);
}

// FFTree::redc_impl, src/fftree.rs:232-259, for nvec vectors of length len sharing `a`
// (plain form).  a0inv (= 1/a[2i], plain) may be supplied when already known.
void Engine::redc(const Fp* evals, const Fp* a_plain, const Fp* a0inv_or_null, size_t len, size_t nvec, Moiety moiety, Fp* out,
                  const Fp* c_or_null, Fp* const* tabs_or_null, const ExitSplit* split, const RedcChain* chain) const {
  const Level& lv = level_for(len);
  if (len < 2) throw Error(ERR_INVALID_ARG, "redc: length must be >= 2");
  const Fp* zinv = moiety == S0 ? lv.z0_inv_s1 : lv.z1_inv_s0;
  if (!zinv) throw Error(ERR_MISSING_TABLES, "redc: tree was built without the Z tables");
  const size_t h = len / 2;
  const uint32_t log_h = ilog2(h);
  const Moiety other = moiety == S1 ? S0 : S1;
  Fp* a0inv_own = nullptr;
  if (!a0inv_or_null && !tabs_or_null) {
    a0inv_own = tmp(h);
    k::copy_strided(a0inv_own, a_plain, h, 2, st);
    k::batch_inverse(a0inv_own, h, st);
    a0inv_or_null = a0inv_own;
  }
  static const bool no_fuse = getenv("ECFFT_B200_NO_REDC_FUSION") != nullptr;
  if (tabs_or_null && !(lv.sym && lv.has_norm() && k::butterfly_mode() != 0 && !no_fuse && log_h >= 1 && (nvec << log_h) >= 4))
    throw Error(ERR_INVALID_ARG, "redc: prebuilt tables need the fused form");
  if (lv.sym && lv.has_norm() && k::butterfly_mode() != 0 && !no_fuse && log_h >= 1 && (nvec << log_h) >= 4) {
    // Fused form: the de-interleave, the two pointwise steps and the interleave ride the two EXTENDs as
    // stride-2 views and per-position tables (k_extend_sym): no pointwise pass over the data.
    //   EXTEND 1: reads evals[2i], pre-scale a0inv*gami (*c), stores out[2i+1] = evals[2i+1]*zinv(*c) - g1^*(gam*a*zinv)
    //   EXTEND 2: reads out[2i+1] (= h1), stores out[2i] = h0
    Fp* tabs = tabs_or_null ? nullptr : tmp(3 * h);
    Fp* work = tmp(h * nvec);
    Fp *P1 = tabs_or_null ? tabs_or_null[0] : tabs, *Kp = tabs_or_null ? tabs_or_null[1] : tabs + h, *Zc = tabs_or_null ? tabs_or_null[2] : tabs + 2 * h;
    if (!tabs_or_null) k::redc_tables(P1, Kp, Zc, a_plain, a0inv_or_null, zinv, lv.gami[moiety], lv.gam[other], c_or_null, h, st);
    k::SymIO io1{1, 0, 1, 1, evals, 1, 1, Zc, work};
    k::SymIO io2{1, 1, 1, 0, nullptr, 0, 0, nullptr, work};
    if (split) {   // EXIT: the second EXTEND stores [u0 | (e0 - u0) * xnn_inv] into the next depth's array instead of out[2i]
      io2.out_shift = 0; io2.out_off = 0;
      io2.E = split->evals; io2.e_shift = 1; io2.e_off = 0; io2.Z = split->xinv_even;
      io2.split = 1;
    }
    const Fp* pre1 = chain && chain->pre_applied ? nullptr : P1;
    const Fp* post2 = chain && chain->post_override ? chain->post_override : lv.gam[moiety];
    const bool ok1 = k::extend_sym(lv.tw_d[moiety], lv.tw_r[other], lv.ctr[other], evals, out, log_h, nvec, pre1, Kp, nullptr, st, &io1);
    const bool ok2 = ok1 && k::extend_sym(lv.tw_d[other], lv.tw_r[moiety], lv.ctr[moiety], out, split ? split->next : out, log_h, nvec, lv.gami[other], post2, nullptr, st, &io2);
    release(tabs);
    release(work);
    release(a0inv_own);
    if (!ok2) throw Error(ERR_INVALID_ARG, "redc: fused EXTEND refused an input it should take");
    return;
  }
  const Fp* src = evals;
  Fp* scaled = nullptr;
  if (c_or_null) {
    scaled = tmp(len * nvec);
    k::mul_bcast(scaled, evals, c_or_null, len, nvec, st);
    src = scaled;
  }
  Fp* t0 = tmp(h * nvec);
  Fp* g1 = tmp(h * nvec);
  k::redc_pre(t0, src, a0inv_or_null, h, nvec, st);
  k::extend(lv, t0, g1, log_h, nvec, other, st);
  Fp* h1 = t0;  // t0 is dead once g1 exists
  k::redc_mid(h1, src, g1, a_plain, zinv, h, nvec, st);
  Fp* h0 = g1;
  k::extend(lv, h1, h0, log_h, nvec, moiety, st);
  k::interleave(out, h0, h1, h * nvec, st);
  release(t0);
  release(g1);
  release(scaled);
  release(a0inv_own);
}

// FFTree::modular_reduce_impl, src/fftree.rs:277-281
void Engine::modular_reduce(const Fp* evals, const Fp* a_plain, const Fp* a0inv_or_null, const Fp* c_plain, size_t len, size_t nvec, Fp* out) const {
  const size_t h = len / 2;
  Fp* a0inv_own = nullptr;
  if (!a0inv_or_null) {
    a0inv_own = tmp(h);
    k::copy_strided(a0inv_own, a_plain, h, 2, st);
    k::batch_inverse(a0inv_own, h, st);
    a0inv_or_null = a0inv_own;
  }
  Fp* hb = tmp(len * nvec);
  redc(evals, a_plain, a0inv_or_null, len, nvec, S0, hb);
  redc(hb, a_plain, a0inv_or_null, len, nvec, S0, out, c_plain);  // the multiplication by c rides the second REDC
  release(hb);
  release(a0inv_own);
}

// FFTree::exit_impl, src/fftree.rs:200-224, flattened top-down: before the pass for m the array
// holds n/m evaluation vectors of length m; the pass replaces each by [u0 | v0].
void Engine::exit(const Fp* evals, Fp* out, size_t n) const {
  level_for(n);
  if (n == 1) {
    if (evals != out) ECFFT_CUDA(cudaMemcpyAsync(out, evals, sizeof(Fp), cudaMemcpyDeviceToDevice, st));
    return;
  }
  Fp* cur = tmp(n);
  Fp* nxt = tmp(n);
  Fp* M = tmp(n);
  ECFFT_CUDA(cudaMemcpyAsync(cur, evals, n * sizeof(Fp), cudaMemcpyDeviceToDevice, st));
  // After the pass for block size m the array holds independent vectors of length m/2, so once the first
  // depth is done the two halves of the array run on two streams (their kernels fill each other's
  // end-of-launch drain, as in enter_range).
  const bool fork = enter_streams() > 1 && !prof::enabled() && n >= ((size_t)1 << 20);
  exit_depths(cur, nxt, M, n, n, fork ? n / 2 : 1);
  if (fork) {
    const size_t part = n / 2;
    std::swap(cur, nxt);  // exit_depths left the result of its single pass in the second buffer
    ScopedEvent ev_fork, ev_join;
    ECFFT_CUDA(cudaEventRecord(ev_fork.e, st));
    cudaStream_t as = t.aux_stream(0);
    ECFFT_CUDA(cudaStreamWaitEvent(as, ev_fork.e, 0));
    Engine sub(t, as);
    sub.exit_depths(cur + part, nxt + part, M + part, part, part, 1);
    ECFFT_CUDA(cudaEventRecord(ev_join.e, as));
    const bool odd_b = exit_depths(cur, nxt, M, part, part, 1);
    ECFFT_CUDA(cudaStreamWaitEvent(st, ev_join.e, 0));
    if (odd_b) std::swap(cur, nxt);  // both halves run the same number of passes
  } else if (ilog2(n) & 1) {
    std::swap(cur, nxt);
  }
  ECFFT_CUDA(cudaMemcpyAsync(out, cur, n * sizeof(Fp), cudaMemcpyDeviceToDevice, st));
  release(cur);
  release(nxt);
  release(M);
}

// The fused-REDC tables of EXIT's MOD at this level (a = xnn_s, c = z0z0_rem_xnn_s), built once per level.
bool Engine::exit_tabs(const Level& lv) const {
  static const bool no_fuse = getenv("ECFFT_B200_NO_REDC_FUSION") != nullptr || getenv("ECFFT_B200_NO_EXIT_TABLES") != nullptr;
  const size_t h = (size_t)1 << (lv.log_n - 1);
  if (no_fuse || !(lv.sym && lv.has_norm() && k::butterfly_mode() != 0) || lv.log_n < 2) return false;
  std::lock_guard<std::mutex> lock(t.tab_mu);
  if (lv.exit_tab[1][2]) return true;
  // built on the tree's own stream and completed before anybody uses them (one-time cost per level)
  cudaStream_t bs = t.stream;
  // stream-ordered allocation: cudaMalloc may synchronise the whole device, which must not happen while another
  // rank's stream of this process waits for work this thread has yet to enqueue (virtual ranks on one GPU)
  auto alloc = [&](size_t count) {
    void* p = nullptr;
    ECFFT_CUDA(cudaMallocAsync(&p, count * sizeof(Fp), bs));
    t.owned_lazy.push_back(p);
    return (Fp*)p;
  };
  Fp* a0inv = alloc(h);
  k::copy_strided(a0inv, lv.xnn_s_inv, h, 2, bs);   // the reference batch-inverts xnn_s[::2] on every call (fftree.rs:235): same values
  Fp* tab[2][3];
  for (int r = 0; r < 2; r++) {
    for (int i = 0; i < 3; i++) tab[r][i] = alloc(h);
    k::redc_tables(tab[r][0], tab[r][1], tab[r][2], lv.xnn_s, a0inv, lv.z0_inv_s1, lv.gami[S0], lv.gam[S1], r ? lv.z0z0 : nullptr, h, bs);
  }
  Fp* post1 = alloc(h);
  k::mul_strided(post1, lv.gam[S0], tab[1][0], 1, 0, h, bs);
  ECFFT_CUDA(cudaStreamSynchronize(bs));
  lv.exit_a0inv = a0inv;
  lv.exit_post1 = post1;
  for (int r = 0; r < 2; r++)
    for (int i = 0; i < 3; i++) lv.exit_tab[r][i] = tab[r][i];
  return true;
}

// The passes of EXIT for block sizes m_from, m_from/2, ..., 2*m_stop on an array of len elements (len/m
// vectors per pass), ping-ponging between cur and nxt.  Returns true when the result ended up in nxt.
bool Engine::exit_depths(Fp* cur, Fp* nxt, Fp* M, size_t len, size_t m_from, size_t m_stop) const {
  bool in_nxt = false;
  for (size_t m = m_from; m >= 2 && m > m_stop; m /= 2) {
    const Level& lv = level_for(m);
    if (!lv.z0z0) throw Error(ERR_MISSING_TABLES, "exit: tree was built without the Z tables");
    const size_t h = m / 2, nvec = len / m;
    if (h >= 2 && (nvec * h) >= 4 && exit_tabs(lv)) {
      // MOD = REDC, x c, REDC (fftree.rs:277-281) with the level's prebuilt tables: four EXTENDs, no pointwise pass
      Fp* hb = tmp(m * nvec);
      // ECFFT_B200_NO_EXIT_CHAIN: REDC 2 applies its own pre-scale (A/B switch; chained: one product per element and depth fewer)
      static const bool no_chain = getenv("ECFFT_B200_NO_EXIT_CHAIN") != nullptr;
      RedcChain c1, c2;
      if (!no_chain) {
        c1.post_override = lv.exit_post1;
        c2.pre_applied = true;
      }
      redc(cur, lv.xnn_s, lv.exit_a0inv, m, nvec, S0, hb, nullptr, lv.exit_tab[0], nullptr, &c1);
      static const bool no_split = getenv("ECFFT_B200_NO_EXIT_SPLIT_FUSION") != nullptr;
      if (!no_split) {
        ExitSplit sp{cur, lv.exit_a0inv, nxt};
        redc(hb, lv.xnn_s, lv.exit_a0inv, m, nvec, S0, M, lv.z0z0, lv.exit_tab[1], &sp, &c2);   // ... and the split rides its last EXTEND
        release(hb);
        std::swap(cur, nxt);
        in_nxt = !in_nxt;
        continue;
      }
      redc(hb, lv.xnn_s, lv.exit_a0inv, m, nvec, S0, M, lv.z0z0, lv.exit_tab[1], nullptr, &c2);
      release(hb);
    } else {
      // the reference batch-inverts xnn_s[::2] on every call (fftree.rs:235); the stored xnn_s_inv holds the same values
      Fp* a0inv = tmp(h);
      k::copy_strided(a0inv, lv.xnn_s_inv, h, 2, st);
      modular_reduce(cur, lv.xnn_s, a0inv, lv.z0z0, m, nvec, M);
      release(a0inv);
    }
    k::exit_split(nxt, cur, M, lv.xnn_s_inv, h, nvec, st);
    std::swap(cur, nxt);
    in_nxt = !in_nxt;
  }
  return in_nxt;
}

// FFTree::mextend_impl, src/fftree.rs:128-135
void Engine::mextend(const Fp* in, Fp* out, size_t h, Moiety target, DataForm form) const {
  const Level& lv = level_for(h * 2);
  const Fp* z = target == S1 ? lv.z0_s1 : lv.z1_s0;
  if (!z) throw Error(ERR_MISSING_TABLES, "mextend: tree was built without the Z tables");
  k::extend(lv, in, out, ilog2(h), 1, target, st);
  k::add_bcast_scaled(out, out, z, form == FORM_MONT ? fp_const_R() : fp_one(), h, 1, st);
}

// FFTree::degree_impl, src/fftree.rs:169-192.  The recursion follows ONE data-dependent branch per level (g1 == e1?).
// With the symmetric tables the branch is taken on the device: the level's difference count stays in device memory,
// one pointwise kernel does either side's pointwise step, and the second EXTEND is launched with that count as its
// condition (k_extend_sym returns at once when it is zero) — no host round trip per level; the result is read back
// once.  The last levels (fewer than 16 elements) and trees without those tables read the count back per level.
size_t Engine::degree(const Fp* evals, size_t n) const {
  level_for(n);
  if (n == 1) return 0;
  Fp* e0 = tmp(n / 2);
  Fp* e1 = tmp(n / 2);
  Fp* g1 = tmp(n / 2);
  Fp* curbuf = tmp(n);
  Fp* work = tmp(n / 2);
  unsigned long long* counters = nullptr;   // [level] difference counts, [64] the degree accumulated on the device
  ECFFT_CUDA(cudaMallocAsync((void**)&counters, 65 * sizeof(unsigned long long), st));
  ECFFT_CUDA(cudaMemsetAsync(counters, 0, 65 * sizeof(unsigned long long), st));
  static const bool host_branch = getenv("ECFFT_B200_DEGREE_HOST_BRANCH") != nullptr;
  const Fp* cur = evals;
  size_t len = n, result = 0;
  uint32_t lvl = 0;
  bool on_device = false;
  while (len > 1) {
    const Level& lv = level_for(len);
    if (!lv.z0_inv_s1) throw Error(ERR_MISSING_TABLES, "degree: tree was built without the Z tables");
    const size_t h = len / 2;
    const uint32_t log_h = ilog2(h);
    k::deinterleave(e0, e1, cur, h, st);
    k::extend(lv, e0, g1, log_h, 1, S1, st);
    k::count_neq(counters + lvl, g1, e1, h, st);
    if (!host_branch && h >= 8 && lv.sym && lv.has_norm() && k::butterfly_mode() == 2) {
      k::degree_step(counters + lvl, e1, g1, lv.z0_inv_s1, e0, curbuf, h, counters + 64, st);
      k::SymIO io{0, 0, 0, 0, nullptr, 0, 0, nullptr, work};
      io.cond = counters + lvl;
      if (!k::extend_sym(lv.tw_d[S1], lv.tw_r[S0], lv.ctr[S0], e1, curbuf, log_h, 1, lv.gami[S1], lv.gam[S0], nullptr, st, &io))
        throw Error(ERR_INVALID_ARG, "degree: EXTEND refused an input it should take");
      on_device = true;
    } else {
      unsigned long long diff = 0;
      ECFFT_CUDA(cudaMemcpyAsync(&diff, counters + lvl, sizeof diff, cudaMemcpyDeviceToHost, st));
      ECFFT_CUDA(cudaStreamSynchronize(st));
      if (diff == 0) {
        ECFFT_CUDA(cudaMemcpyAsync(curbuf, e0, h * sizeof(Fp), cudaMemcpyDeviceToDevice, st));
      } else {
        k::sub_mul_bcast(e1, e1, g1, lv.z0_inv_s1, h, 1, st);  // t1
        k::extend(lv, e1, curbuf, log_h, 1, S0, st);            // t0
        result += h;
      }
    }
    cur = curbuf;
    len = h;
    lvl++;
  }
  if (on_device) {
    unsigned long long dev_part = 0;
    ECFFT_CUDA(cudaMemcpyAsync(&dev_part, counters + 64, sizeof dev_part, cudaMemcpyDeviceToHost, st));
    ECFFT_CUDA(cudaStreamSynchronize(st));
    result += (size_t)dev_part;
  }
  release(e0);
  release(e1);
  release(g1);
  release(curbuf);
  release(work);
  ECFFT_CUDA(cudaFreeAsync(counters, st));
  return result;
}

void Engine::redc_user(const Fp* evals, const Fp* a_mont, size_t n, Moiety moiety, Fp* out) const {
  level_for(n);
  Fp* a_plain = tmp(n);
  k::mul_const(a_plain, a_mont, fp_const_RINV(), n, st);
  redc(evals, a_plain, nullptr, n, 1, moiety, out);
  release(a_plain);
}
void Engine::mod_user(const Fp* evals, const Fp* a_mont, const Fp* c_mont, size_t n, Fp* out) const {
  level_for(n);
  Fp* a_plain = tmp(n);
  Fp* c_plain = tmp(n);
  k::mul_const(a_plain, a_mont, fp_const_RINV(), n, st);
  k::mul_const(c_plain, c_mont, fp_const_RINV(), n, st);
  modular_reduce(evals, a_plain, nullptr, c_plain, n, 1, out);
  release(a_plain);
  release(c_plain);
}

// FFTree::vanish_impl, src/fftree.rs:291-308, bottom-up.  Products of two data values need the
// plain domain, so Montgomery-form input is converted on entry and the result on exit.
void Engine::vanish(const Fp* domain, Fp* out, size_t n, DataForm form) const {
  level_for(2 * n);
  if (t.log_n < 1) throw Error(ERR_TREE_TOO_SMALL, "FFTree is too small");
  Fp* Q = tmp(2 * n);
  Fp* Q2 = tmp(2 * n);
  Fp* q0 = tmp(n);
  Fp* e = tmp(n);
  const Fp* dom = domain;
  if (form == FORM_MONT) {
    k::mul_const(e, domain, fp_const_RINV(), n, st);
    dom = e;
  }
  // base case: the 2-leaf tree's leaves are f[n_top], f[n_top + n_top/2]
  k::vanish_base(Q, dom, t.base_leaf0, t.base_leaf1, n, st);
  for (size_t len = 2, cnt = n; cnt > 1; len *= 2, cnt /= 2) {
    const Level& lv = level_for(2 * len);
    if (!lv.z0_s1) throw Error(ERR_MISSING_TABLES, "vanish: tree was built without the Z tables");
    const size_t pairs = cnt / 2;
    static const bool no_fuse = getenv("ECFFT_B200_NO_VANISH_FUSION") != nullptr;
    if (!no_fuse && lv.sym && lv.has_norm() && k::butterfly_mode() == 2 && len >= 2 && len * pairs >= 4) {
      // Fused form: the sibling products go straight to the even slots of the next array; MEXTEND reads them there as
      // a stride-2 view and stores e + Z_0 into the odd slots (k_extend_sym's store epilogue): one pointwise pass and
      // the EXTEND per depth instead of three passes and the EXTEND.
      k::mul_pairs_even(Q2, Q, len, pairs, st);
      k::SymIO io{1, 0, 1, 1, nullptr, 0, 0, lv.z0_s1, e};
      if (!k::extend_sym(lv.tw_d[S0], lv.tw_r[S1], lv.ctr[S1], Q2, Q2, ilog2(len), pairs, lv.gami[S0], lv.gam[S1], nullptr, st, &io))
        throw Error(ERR_INVALID_ARG, "vanish: fused MEXTEND refused an input it should take");
      Fp* sw = Q;
      Q = Q2;
      Q2 = sw;
      continue;
    }
    k::mul_pairs(q0, Q, len, pairs, 0, st);
    k::extend(lv, q0, e, ilog2(len), pairs, S1, st);
    k::vanish_merge(Q2, q0, e, lv.z0_s1, fp_one(), len, pairs, st);
    Fp* sw = Q;
    Q = Q2;
    Q2 = sw;
  }
  if (form == FORM_MONT)
    k::mul_const(out, Q, fp_const_R(), 2 * n, st);
  else
    ECFFT_CUDA(cudaMemcpyAsync(out, Q, 2 * n * sizeof(Fp), cudaMemcpyDeviceToDevice, st));
  release(Q);
  release(Q2);
  release(q0);
  release(e);
}

}  // namespace ecfft
