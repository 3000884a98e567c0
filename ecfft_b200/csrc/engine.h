// Internal interface of the B200 ECFFT engine (not installed; the public boundary is
// include/ecfft_b200.h).  Everything here works on DEVICE pointers to Fp arrays.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "fp.cuh"

namespace ecfft {

// status codes mirrored in include/ecfft_b200.h
enum Status : int {
  OK = 0,
  ERR_NOT_POW2 = 1,        // reference: assert!(n.is_power_of_two()) fftree.rs:44,490
  ERR_TREE_TOO_SMALL = 2,  // reference: panic!("FFTree is too small") fftree.rs:494
  ERR_BAD_BYTES = 3,       // reference: SerializationError (truncated / element >= p)
  ERR_CUDA = 4,
  ERR_INVALID_ARG = 5,
  ERR_TOO_LARGE = 6,       // reference: build_fftree returns None when log2 n >= 36, lib.rs:61-64
  ERR_MISSING_TABLES = 7,  // tree built with ENTER-only tables asked for EXIT/REDC/...
  ERR_BUFFER_TOO_SMALL = 8,
};

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

#define ECFFT_CUDA(expr)                                                                          \
  do {                                                                                            \
    cudaError_t e__ = (expr);                                                                     \
    if (e__ != cudaSuccess)                                                                       \
      throw ::ecfft::Error(::ecfft::ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
  } while (0)

// Every entry point works on the tree's device and leaves the caller's current device as it found it.
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int device) {
    if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
    if (prev != device) ECFFT_CUDA(cudaSetDevice(device));
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
  DeviceGuard(const DeviceGuard&) = delete;
  DeviceGuard& operator=(const DeviceGuard&) = delete;
};

// cudaFuncSetAttribute is per device: a mutex-guarded bitmask of the devices a kernel has been configured on
struct PerDeviceOnce {
  std::mutex mu;
  unsigned long long done = 0;
  template <class F>
  void run(F f) {
    int d = 0;
    ECFFT_CUDA(cudaGetDevice(&d));
    std::lock_guard<std::mutex> lock(mu);
    if (d < 64 && ((done >> d) & 1)) return;
    f();
    if (d < 64) done |= 1ull << d;
  }
};

enum Moiety : int { S0 = 0, S1 = 1 };  // reference fftree.rs:17-21, declaration order
// Whether the DATA flowing through an op is in Montgomery form (public API) or plain form
// (tree construction).  Tables are always plain.  Only additive constants and
// data*data products care.
enum DataForm : int { FORM_MONT = 0, FORM_PLAIN = 1 };

// One level of the subtree chain: FFTree with N = 2^log_n leaves (reference fftree.rs:23-38).
// All tables are device arrays of canonical plain-form Fp.  Matrices keep the reference's
// BinaryTree layout: N entries of 4 Fp (row major), layer l at [ (N/2)>>l , 2*((N/2)>>l) ).
struct Level {
  uint32_t log_n = 0;
  Fp* rmat = nullptr;       // N * 4
  Fp* dmat = nullptr;       // N * 4
  Fp* xnn_s = nullptr;      // N
  Fp* xnn_s_inv = nullptr;  // N
  Fp* z0_s1 = nullptr;      // N/2
  Fp* z1_s0 = nullptr;
  Fp* z0_inv_s1 = nullptr;
  Fp* z1_inv_s0 = nullptr;
  Fp* z0z0 = nullptr;       // N   <Z_0^2 mod X^(N/2) on S>
  Fp* z1z1 = nullptr;       // N
  bool has_z = false;       // z tables present (full build)
  // Normalised-butterfly tables (DESIGN.md "twiddle form"), derived from f and rmat; h = N/2.
  // Index (2^j + i) addresses butterfly i of the level with half-stride 2^j; moiety mu = 0/1.
  // sym = false ("normalised", 2 products per pair): tw_r = {s0, s1} (the two tree nodes of the pair),
  //   tw_d = {-s1, -s0}, 2 Fp per butterfly; gami also carries the sum-form input scalings.
  // sym = true ("symmetric", 1 product per pair; maps of the form (x^2 + c1 x + beta^2)/x):
  //   tw_r = g = (s0 - beta)/(s0 + beta), tw_d = 1/g, 1 Fp per butterfly; gam = prod_j (s + beta_j) v_j(s),
  //   gami = 2^-L / gam.
  bool sym = false;
  Fp* tw_r[2] = {nullptr, nullptr};
  Fp* tw_d[2] = {nullptr, nullptr};
  Fp* gam[2] = {nullptr, nullptr};    // h: accumulated scale Gamma^mu_p of the recombine network
  Fp* gami[2] = {nullptr, nullptr};   // h: pre-scale of the decompose network (1/Gamma^mu_p and folded constants)
  Fp* gx = nullptr;                   // h: Gamma^1_i * xnn_s[2i+1] (ENTER combine)
  Fp* ctr[2] = {nullptr, nullptr};    // sym: 1 element, g_target/g_source at level 0 (index = target moiety)
  // EXIT's two REDCs per depth (fftree.rs:206-210) always use a = xnn_s, c = z0z0_rem_xnn_s: their fused-REDC
  // tables (kernels.cu redc_tables) depend on the level only and are built once, on first use (Engine::exit_tabs)
  mutable Fp* exit_tab[2][3] = {{nullptr, nullptr, nullptr}, {nullptr, nullptr, nullptr}};  // [redc 0 | redc 1 (x c)][P1, Kp, Zc], h each
  mutable Fp* exit_a0inv = nullptr;   // xnn_s_inv[::2], h
  mutable Fp* exit_post1 = nullptr;   // gam[S0] * exit_tab[1][0]: REDC 1's last post-scale with REDC 2's first pre-scale folded in, h
  // ENTER with the next depth's pre-scale folded into this depth's combine (Engine::fold_tabs; built on first use, h each).
  // P = gami[0] of this level (the scale its input carries when "folded"), Pn = gami[0] of the level above (the scale
  // its output is to carry):  [0] gam[1][i] Pn[2i+1], [1] gx[i] Pn[2i+1]              (odd outputs, folded out)
  //   [2] Pn[2i] / P[i],  [3] xnn[2i] Pn[2i] / P[i]   (even outputs, folded in and out)
  //   [4] 1 / P[i],       [5] xnn[2i] / P[i]          (folded in, plain out)
  //   [6] Pn[2i],         [7] xnn[2i] Pn[2i]          (plain in, folded out)
  mutable Fp* fold_tab[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  bool has_norm() const { return tw_r[0] && tw_r[1] && tw_d[0] && tw_d[1] && gam[0] && gam[1] && gami[0] && gami[1] && gx; }
};

struct RatMapHost {  // reference utils.rs:367-371; coefficients low->high, plain canonical
  std::vector<Fp> num, den;
};

enum BuildParts : int { PARTS_FULL = 0, PARTS_ENTER_ONLY = 1 };

struct Tree {
  int device = 0;
  uint32_t log_n = 0;                // top level
  Fp* f = nullptr;                   // top-level BinaryTree<F>, 2n entries (f[0] = 0)
  std::vector<Level> levels;         // levels[k] has 2^k leaves, k = 0..log_n
  std::vector<RatMapHost> maps;      // log_n rational maps (level k uses the first k)
  std::vector<Fp> beta;              // per map: fixed point of its deck involution x -> beta^2/x (symmetric butterflies)
  bool sym_ok = false;               // every map is (x^2 + c1 x + beta^2)/x
  int parts = PARTS_FULL;
  Fp base_leaf0, base_leaf1;         // leaves of the 2-leaf chain level (VANISH base case, fftree.rs:293-298)
  unsigned long long* build_errors = nullptr;  // device counter: zero denominators / determinants / twiddles met while building
  cudaStream_t stream = nullptr;     // default stream for host-buffer calls
  std::vector<void*> owned;          // device allocations to free
  mutable std::vector<cudaStream_t> aux;  // helper streams ENTER forks independent ranges onto (engine.cu)
  mutable std::mutex aux_mu;
  mutable std::mutex tab_mu;             // guards the lazily built per-level EXIT tables
  mutable std::vector<void*> owned_lazy;  // ... and owns them
  cudaStream_t aux_stream(int i) const;
  ~Tree();
  Fp* dalloc(size_t count);          // owned device allocation of `count` Fp
  size_t n() const { return (size_t)1 << log_n; }
};

// ---- instrumentation (bench.py): launch counter and optional per-launch CUDA-event timing of the
// two hot kernels, with the ALGORITHMIC bytes (level-streaming model) each launch covers
namespace prof {
enum Kernel : int { EXTEND_TILE = 0, ENTER_COMBINE = 1, NUM_KERNELS = 2 };
void count_launch();
unsigned long long launches();
void enable(bool on);
bool enabled();
void record_begin(Kernel k, double alg_bytes, cudaStream_t st);
void record_end(cudaStream_t st);
void read(Kernel k, double* ms, double* alg_bytes, unsigned long long* launches);  // synchronises; clears k's records
}  // namespace prof

// ---- kernels.cu: launchers (all asynchronous on `st`) -------------------------------------
namespace k {
// EXTEND of `nvec` contiguous vectors of length h = 2^log_h on the level with 2h leaves,
// towards `target`; in may equal out.  With the normalised tables present (and
// butterfly_mode() == 1) it runs the 2-multiplication butterflies; `unscaled_out` then leaves
// the final Gamma scaling to the caller (ENTER folds it into its combine).
void extend(const Level& lv, const Fp* in, Fp* out, uint32_t log_h, size_t nvec, Moiety target, cudaStream_t st,
            bool unscaled_out = false);
void extend_sub(const Level& lv, const Fp* in, Fp* out, uint32_t log_len, cudaStream_t st, const Fp* pre = nullptr, Moiety source = S0, Moiety target = S1);
// sym_kernel.cu: all passes of the symmetric-butterfly EXTEND (tw_d = 1/g of the source moiety, tw_r = g of
// the target moiety, ctr = Level::ctr[target]; pre/post = per-position scales or null).  comb != null fuses ENTER's combine
// (fftree.rs:155-159) into the last pass: vectors 2w, 2w+1 are u, v of block w, comb->A the depth's
// unscaled input, and the result goes to comb->out.  Returns false when it does not apply (fewer than 4
// elements; a depth whose vector length equals the tile cannot take the fused combine).
// e0 != null: the even output is e0[i] u0 + e1[i] v0 (one reduction) instead of u0 + v0 xnn[2i] — the folded forms.
struct SymCombine { const Fp* A; const Fp* xnn; const Fp* gam; const Fp* gx; Fp* out; const Fp* e0 = nullptr; const Fp* e1 = nullptr; };
// One pass of k_extend_sym (a launch of the per-pass kernel, or one entry of a flow: sym_kernel.cu)
struct SymParams {
  const Fp* in;
  Fp* out;
  const Fp* tw_d;   // 1/g of the source moiety, entry 2^j + i
  const Fp* tw_r;   // g of the target moiety
  const Fp* ctr;    // one element: g_target / g_source at level 0 (the centre of the network)
  const Fp* pre;    // per-position scale applied by the first stage (or null)
  const Fp* post;   // per-position scale applied by the last stage (or null); ignored when comb != 0
  const Fp* A;      // combine epilogue: the unscaled input vectors [u0 | v0] per block
  const Fp* xnn;
  const Fp* gam;
  const Fp* gx;
  const Fp* ce0;    // combine epilogue, folded forms: even output = ce0[i] u0 + ce1[i] v0 (null: u0 + v0 xnn[2i])
  const Fp* ce1;
  unsigned long long nv;      // strided: vectors (pair: vector pairs) in the batch; blocks are ordered batch-major
  unsigned long long total;   // elements in the batch (guards the ragged tile of tiny inputs)
  uint32_t log_h, log_t;
  uint32_t lvl_lo, lvl_hi;    // this pass runs the levels lvl_lo <= j < lvl_hi
  uint32_t boff;              // tile-index bit of level j is j + boff (mod 2^32)
  uint32_t packed;            // 1: tile = 2^log_t consecutive elements (whole vectors, or a slice of one: inner pass)
  uint32_t log_c, krows, row_shift;  // strided: 2^krows rows of 2^log_c contiguous elements, rows 2^row_shift apart
  uint32_t pair;              // strided: top tile bit selects vector 2w / 2w+1
  uint32_t comb;              // 1: combine epilogue
  uint32_t do_d, do_r;
  // strided views (REDC's de-interleave / interleave, src/fftree.rs:234, 258): logical element g of the
  // batch is read at in[(g << in_shift) + in_off] by the first pass and written at out[(g << out_shift) +
  // out_off] by the last; E/Z: the last pass stores E[(g << e_shift) + e_off] * Z[i] + x * post[i]
  uint32_t in_shift, in_off, out_shift, out_off, e_shift, e_off;
  const Fp* E;
  const Fp* Z;
  // split != 0 (EXIT's last EXTEND of a depth, src/fftree.rs:206-220): u0 = x * post is stored at
  // out[(vector << (log_h + 1)) + i] and v0 = (E[(g << e_shift) + e_off] - u0) * Z[i] at the same place + h
  uint32_t split;
  const unsigned long long* cond;   // not null: the whole pass is skipped when *cond == 0 (DEGREE's data-dependent branch, on the device)
  uint32_t tma_fence;
  uint32_t l2pf;              // strided tiles: bulk L2 prefetch of the tile's twiddle / combine-table ranges (set by launch_sym)
  uint32_t pf;                // twiddle prefetch ahead of each stage: 0 off, 1 into L1, 2 into L2 (set by launch_sym)
  // ---- flow fields (k_sym_flow: all passes of an ENTER in one persistent launch) ----
  uint32_t kind;              // 0: butterfly tile pass; 1: combine-only pass (in = [u1 | v1] unscaled, A = [u0 | v0])
  uint32_t tile_begin, ntiles;         // this pass's range of the flow's tile queue
  uint32_t ord_tpb_log, ord_nv, ord_log_v;  // queue order within the pass: (block, vector, tile in block), see finalize_flow
  uint32_t dep_base, dep_shift, dep_need;   // wait until counter[dep_base + (first input element >> dep_shift)] >= dep_need (0: none)
  uint32_t sig_base, sig_shift;             // then add 1 to counter[sig_base + (first output element >> sig_shift)]
};
// A flow: passes whose tiles depend only on an aligned block of the previous pass's output (butterfly networks
// and ENTER's combine are block-local), executed by ONE launch of persistent CTAs that take tiles from a queue
// in dependency order and wait on per-block completion counters instead of kernel boundaries.
struct SymFlow {
  std::vector<SymParams> passes;
  double alg_bytes = 0;   // level-streaming model, summed (bench roofline)
};
// appends the passes of one symmetric EXTEND (same arguments as extend_sym) / one combine-only pass; false: not applicable
bool plan_extend_sym(SymFlow& flow, const Fp* tw_d, const Fp* tw_r, const Fp* ctr, const Fp* in, Fp* out, uint32_t log_h, size_t nvec,
                     const Fp* pre, const Fp* post, const SymCombine* comb, const struct SymIO* io = nullptr);
void plan_combine_only(SymFlow& flow, const SymCombine& c, const Fp* W, uint32_t log_h, size_t n);
bool flow_enabled();            // ECFFT_B200_FLOW (default 1)
uint32_t flow_log_tile();
void launch_flow(SymFlow& flow, cudaStream_t st);   // throws if the flow does not fit one launch
void flow_stats_enable(bool on);                     // diagnostics: per-launch wait / body / signal cycles of the flow CTAs
void flow_stats_read(unsigned long long out4[4]);    // {wait cycles, body cycles, signal cycles, tiles}; clears
// Strided views for REDC (fftree.rs:232-259): logical element g is read at in[(g << in_shift) + in_off] and
// written at out[(g << out_shift) + out_off]; with E the store is E[(g << e_shift) + e_off] * Z[i] + x * post[i]
// (i = position within the vector).  work: contiguous scratch of nvec * h elements for multi-pass EXTENDs.
struct SymIO { uint32_t in_shift, in_off, out_shift, out_off; const Fp* E; uint32_t e_shift, e_off; const Fp* Z; Fp* work; uint32_t split = 0;
               const unsigned long long* cond = nullptr; };   // cond: every pass is skipped when *cond == 0
bool extend_sym(const Fp* tw_d, const Fp* tw_r, const Fp* ctr, const Fp* in, Fp* out, uint32_t log_h, size_t nvec, const Fp* pre, const Fp* post,
                const SymCombine* comb, cudaStream_t st, const SymIO* io = nullptr);
void mg_cross(const Level& lv, int phase, uint32_t j, int role, size_t p_pos0, const Fp* own, const Fp* partner, size_t count, Fp* out, cudaStream_t st,
              Moiety source = S0, Moiety target = S1);
void dot2_strided(Fp* out, const Fp* e, size_t e_stride, const Fp* z, const Fp* g, const Fp* kp, size_t n, cudaStream_t st);
void sub_mul_strided(Fp* out, const Fp* a, size_t a_stride, const Fp* b, const Fp* c, size_t n, cudaStream_t st);
void mg_sync(unsigned long long* own_flag, unsigned long long value, const unsigned long long* wait_a,
             const unsigned long long* wait_b, unsigned timeout_ms, cudaStream_t st, unsigned long long* status = nullptr, unsigned info = 0);
void mg_wait_all(void* const* bases, int world, int rank, unsigned idx, unsigned long long value, unsigned timeout_ms, cudaStream_t st,
                 unsigned long long* status = nullptr);
void mg_combine(const Level& lv, size_t i0, const Fp* u0, const Fp* v0, const Fp* u1, const Fp* v1, size_t count, Fp* out, cudaStream_t st);
int butterfly_mode();  // ECFFT_B200_BUTTERFLY: 2 = symmetric (default), 1 = normalised, 0 = 2x2 matrices
// ENTER combine, fftree.rs:155-159, batched over n/(2h) blocks.  W_unscaled: W lacks the Gamma^1
// scaling (lv.gam[1], lv.gx are used instead of xnn's odd entries).
void enter_combine(const Level& lv, const Fp* A, const Fp* W, Fp* out, uint32_t log_h, size_t n, bool W_unscaled, cudaStream_t st);
void enter_combine_tabs(const SymCombine& c, const Fp* W, uint32_t log_h, size_t n, cudaStream_t st);  // W unscaled; any table set
// the folded-combine tables of Level::fold_tab: group 0 = [0],[1]; 1 = [2],[3]; 2 = [4],[5]; 3 = [6],[7]
void fold_tables(int group, Fp* t0, Fp* t1, const Fp* gam0, const Fp* gam1, const Fp* gx, const Fp* xnn, const Fp* Pn, Fp two_pow_L, size_t h, cudaStream_t st);

void mul_const(Fp* out, const Fp* in, Fp c, size_t n, cudaStream_t st);                       // out = in*c
void mul_bcast(Fp* out, const Fp* in, const Fp* c, size_t len, size_t nvec, cudaStream_t st); // out[v][i] = in[v][i]*c[i]
void mul_mont(Fp* out, const Fp* a, const Fp* b, size_t n, cudaStream_t st);                    // Montgomery product of two data vectors
void add_bcast_scaled(Fp* out, const Fp* in, const Fp* z, Fp scale, size_t len, size_t nvec, cudaStream_t st);  // out = in + z*scale
void deinterleave(Fp* even, Fp* odd, const Fp* in, size_t pairs, cudaStream_t st);
void interleave(Fp* out, const Fp* even, const Fp* odd, size_t pairs, cudaStream_t st);
void copy_strided(Fp* out, const Fp* in, size_t count, size_t in_stride, cudaStream_t st);   // out[i] = in[i*stride]
// REDC pieces (fftree.rs:232-259), vectors of length 2h, a/zinv shared by all vectors
void redc_pre(Fp* t0, const Fp* evals, const Fp* a0inv, size_t h, size_t nvec, cudaStream_t st);
void redc_mid(Fp* h1, const Fp* evals, const Fp* g1, const Fp* a, const Fp* zinv, size_t h, size_t nvec, cudaStream_t st);
void redc_tables(Fp* P1, Fp* Kp, Fp* Zc, const Fp* a, const Fp* a0inv, const Fp* zinv, const Fp* gami_src, const Fp* gam_tgt,
                 const Fp* c, size_t h, cudaStream_t st);
// EXIT split (fftree.rs:206-220): next[v] = [ M[v][::2] | (e[v][::2]-M[v][::2]) * xnn_inv[::2] ]
void exit_split(Fp* next, const Fp* evals, const Fp* M, const Fp* xnn_inv, size_t h, size_t nvec, cudaStream_t st);
// VANISH pieces (fftree.rs:291-308)
void vanish_base(Fp* out, const Fp* dom, Fp l0, Fp l1, size_t n, cudaStream_t st);            // out[2i] = a-l0, out[2i+1] = a-l1
void mul_pairs(Fp* q0, const Fp* Q, size_t len, size_t npairs, int fix_mont, cudaStream_t st); // q0[w] = Q[2w]*Q[2w+1]
void mul_pairs_even(Fp* out, const Fp* Q, size_t len, size_t npairs, cudaStream_t st);         // out[w][2i] = Q[2w][i]*Q[2w+1][i]
void vanish_merge(Fp* out, const Fp* q0, const Fp* e, const Fp* z, Fp zscale, size_t len, size_t nvec, cudaStream_t st);
// DEGREE pieces (fftree.rs:169-192)
void count_neq(unsigned long long* counter, const Fp* a, const Fp* b, size_t n, cudaStream_t st);
// one level of DEGREE decided on the device (fftree.rs:181-191): *diff == 0: next = e0; else e1 = (e1 - g1) * zinv and *result += h
void degree_step(const unsigned long long* diff, Fp* e1, const Fp* g1, const Fp* zinv, const Fp* e0, Fp* next, size_t h,
                 unsigned long long* result, cudaStream_t st);
void sub_mul_bcast(Fp* out, const Fp* a, const Fp* b, const Fp* c, size_t len, size_t nvec, cudaStream_t st);  // (a-b)*c
// generic helpers
void pow_u64(Fp* out, const Fp* in, uint64_t e, size_t n, cudaStream_t st);
void batch_inverse(Fp* v, size_t n, cudaStream_t st);  // in place; zeros stay zero (ark_ff::batch_inversion)
void count_noncanonical(unsigned long long* counter, const Fp* v, size_t n, cudaStream_t st);
void fill(Fp* out, Fp c, size_t n, cudaStream_t st);
void sqr_sub_mul(Fp* out, const Fp* z_half, int z_parity, const Fp* xnn, const Fp* sub, const Fp* mul, size_t n, cudaStream_t st);
void muladd(Fp* out, const Fp* a, const Fp* b, const Fp* c, size_t n, cudaStream_t st);  // out = a + b*c
// tree construction
void build_leaves(Fp* leaves, size_t n, Fp a, Fp a4, Fp offx, Fp offy, const Fp* gtab_xy /* log n points: 2^j * G */, uint32_t log_n, cudaStream_t st);
// err (may be null): device counter of the inputs the reference would panic on (zero denominator, singular matrix)
void ratmap_layer(Fp* layer, const Fp* prev, size_t count, const Fp* num, int nnum, const Fp* den, int nden, unsigned long long* err, cudaStream_t st);
void build_matrices(Fp* rmat_layer, Fp* dmat_layer, const Fp* flayer, size_t fstride, size_t d, const Fp* den, int nden, unsigned long long* err, cudaStream_t st);
// normalised tables of one chain level (h = N/2 entries each); f_top strided by fstride is the level's f
void build_twiddles(Fp* tw_r, Fp* tw_d, const Fp* f_top, size_t fstride, size_t h, int mu, cudaStream_t st);
void build_gamma(Fp* gam, const Fp* rmat, size_t h, int mu, cudaStream_t st);
void fold_sumform_prescale(Fp* gami, const Fp* f_top, size_t fstride, size_t h, int mu, cudaStream_t st);
// symmetric-form tables; beta_by_j[j] is the fixed point of the map used by the level with half-stride 2^j
void build_twiddles_sym(Fp* tw_r, Fp* tw_d, const Fp* f_top, size_t fstride, size_t h, int mu, const Fp* beta_by_j, unsigned long long* err, cudaStream_t st);
void build_gamma_sym(Fp* gam, const Fp* rmat, const Fp* f_top, size_t fstride, size_t h, int mu, const Fp* beta_by_j, cudaStream_t st);
void mul_strided(Fp* out, const Fp* a, const Fp* b, size_t b_stride, size_t b_off, size_t n, cudaStream_t st);
void selftest_field(unsigned long long* counters3, unsigned long long samples, cudaStream_t st);  // device self-test of the lazy add/sub forms  // out[i] = a[i]*b[b_off + i*b_stride]
}  // namespace k

// ---- engine.cu: the algorithms on device buffers ------------------------------------------
struct Engine {
  const Tree& t;
  cudaStream_t st;
  Engine(const Tree& tree, cudaStream_t s) : t(tree), st(s) {}
  const Level& level_for(size_t leaves) const;  // subtree_with_size; throws like the reference panics

  // scratch (stream-ordered)
  Fp* tmp(size_t count) const;
  void release(Fp* p) const;

  // batched primitives: nvec contiguous vectors
  void extend(const Fp* in, Fp* out, size_t h, size_t nvec, Moiety target) const;
  // c_or_null: evals are to be multiplied pointwise by c first (MOD's middle step, folded into the tables)
  // EXIT (fftree.rs:206-220): the next depth's array [u0 | (e0 - u0) * xnn_inv[::2]] written by REDC's last EXTEND
  struct ExitSplit { const Fp* evals; const Fp* xinv_even; Fp* next; };
  // Two REDCs in a row (MOD inside EXIT): the even outputs of the first feed nothing but the first EXTEND of the second,
  // so its pre-scale rides the first one's last post-scale (post_override) and the second skips it (pre_applied).
  struct RedcChain { const Fp* post_override = nullptr; bool pre_applied = false; };
  void redc(const Fp* evals, const Fp* a_plain, const Fp* a0inv_or_null, size_t len, size_t nvec, Moiety moiety, Fp* out,
            const Fp* c_or_null = nullptr, Fp* const* tabs_or_null = nullptr, const ExitSplit* split = nullptr,
            const RedcChain* chain = nullptr) const;  // tabs: prebuilt {P1, Kp, Zc} of the fused form
  bool exit_tabs(const Level& lv) const;  // builds lv.exit_tab / exit_a0inv once; false: the fused REDC does not apply
  void modular_reduce(const Fp* evals, const Fp* a_plain, const Fp* a0inv_or_null, const Fp* c_plain, size_t len, size_t nvec, Fp* out) const;

  // the FFTree<F> surface (fftree.rs:72-316) on device buffers
  void enter_range(const Fp* in, Fp* out, size_t n, size_t m_lo, size_t m_hi, int streams_hint = 0) const;  // bottom-up levels m_lo < m <= m_hi
  // ... on this engine's stream only.  in_folded / out_folded: the input already carries / the output is to carry the
  // pre-scale of the EXTEND that consumes it next (the data between two depths of one ENTER)
  void enter_range_serial(const Fp* in, Fp* out, size_t n, size_t m_lo, size_t m_hi, bool in_folded = false, bool out_folded = false) const;
  bool fold_tabs(const Level& lv, int group) const;  // builds lv.fold_tab[2 group], [2 group + 1] once
  bool enter_range_flow(const Fp* in, Fp* out, size_t n, size_t m_lo, size_t m_hi) const;    // ... as one flow launch (false: not applicable)
  void enter(const Fp* coeffs, Fp* out, size_t n) const { enter_range(coeffs, out, n, 1, n); }
  void exit(const Fp* evals, Fp* out, size_t n) const;
  bool exit_depths(Fp* cur, Fp* nxt, Fp* M, size_t len, size_t m_from, size_t m_stop) const;
  void mextend(const Fp* in, Fp* out, size_t h, Moiety target, DataForm form) const;
  size_t degree(const Fp* evals, size_t n) const;
  void redc_user(const Fp* evals, const Fp* a_mont, size_t n, Moiety moiety, Fp* out) const;
  void mod_user(const Fp* evals, const Fp* a_mont, const Fp* c_mont, size_t n, Fp* out) const;
  void vanish(const Fp* domain, Fp* out, size_t n, DataForm form) const;
};

// ---- sharded.cu: the per-rank schedule of the multi-GPU ENTER over peer-mapped arenas ----
// Arena of a rank: [ECFFT_MG_FLAG_BYTES of u64 flags | slots of n/world elements]; bases[r] is rank r's
// arena as mapped into this process.  Writes this rank's n/world evaluations (positions
// [rank n/world, (rank+1) n/world)) to out_chunk.
static constexpr size_t MG_FLAG_BYTES = 4096;
static constexpr unsigned MG_DONE_FLAG = MG_FLAG_BYTES / 8 - 1;  // "this rank has finished call `epoch`"
static constexpr unsigned MG_STATUS_FLAG = MG_FLAG_BYTES / 8 - 2;  // 0, or the record of the first wait that timed out
size_t peer_arena_bytes(size_t n, int world);
unsigned peer_timeout_ms();  // ECFFT_B200_PEER_TIMEOUT_MS (default 20000, 0 = wait for ever)
void enter_peer(const Engine& eng, const Fp* chunk, size_t n, int rank, int world, void* const* bases,
                unsigned long long epoch, Fp* out_chunk);
// EXIT (fftree.rs:200-224) the same way: rank `rank` holds evaluations [rank n/world, (rank+1) n/world) and ends
// with coefficients of the same range; the top log2(world) depths run MOD across the ranks, then every rank
// runs an independent EXIT(n/world).  Needs peer_exit_arena_bytes.
size_t peer_exit_arena_bytes(size_t n, int world);
void exit_peer(const Engine& eng, const Fp* chunk, size_t n, int rank, int world, void* const* bases,
               unsigned long long epoch, Fp* out_chunk);

void set_last_error(const char* msg);   // capi.cu: the message ecfft_last_error() returns on this thread

// ---- builder.cu / serialize.cu --------------------------------------------------------------
Tree* build_secp256k1(size_t n, int parts, int device);                                     // lib.rs:39-85
Tree* tree_from_leaves(const Fp* leaves_dev_plain, size_t n, const std::vector<RatMapHost>& maps, int parts, int device);  // fftree.rs:42-70
void finish_tree(Tree& t);
void build_norm_tables(Tree& t, uint32_t level);  // normalised-butterfly tables from f + rmat                                                                  // fftree.rs:318-463 for every chain level
size_t serialized_size(const Tree& t, bool compressed);
size_t serialize(const Tree& t, bool compressed, uint8_t* buf, size_t cap);
Tree* deserialize(const uint8_t* buf, size_t len, bool compressed, int device);

}  // namespace ecfft
