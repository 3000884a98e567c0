// FFTree construction on the GPU.
//   build_secp256k1  — Fp::build_fftree, reference src/lib.rs:39-85 (+ GoodCurve / isogeny chain,
//                      src/ec.rs:37-90,177-189,344-358; two_adicity src/utils.rs:356-365)
//   tree_from_leaves — FFTree::new, src/fftree.rs:42-70
//   finish_tree      — FFTree::from_tree + derive_subtree for every level of the subtree chain,
//                      src/fftree.rs:318-482
// Only O(log^2 n) field operations run on the host (curve constants, isogeny chain); leaves,
// internal nodes, matrices and all evaluation tables are produced by device kernels, the Z
// tables by running EXTEND / VANISH / MOD of this engine on the partially built tree exactly as
// the reference does.  All values are plain (non-Montgomery) canonical integers.
#include <string.h>

#include "ec.cuh"
#include "engine.h"

namespace ecfft {

static inline uint32_t ilog2(size_t n) {
  uint32_t l = 0;
  while (n >>= 1) l++;
  return l;
}

static Fp fp_from_hex(const char* hex) {
  Fp r = fp_zero();
  size_t n = strlen(hex);
  for (size_t i = 0; i < n; i++) {
    char ch = hex[n - 1 - i];
    uint32_t d = (ch >= '0' && ch <= '9') ? ch - '0' : (ch >= 'a' && ch <= 'f') ? ch - 'a' + 10 : ch - 'A' + 10;
    r.v[i / 8] |= d << (4 * (i % 8));
  }
  return r;
}
static bool fp_sqrt_host(const Fp& x, Fp* r) {  // ark-ff sqrt for p = 3 mod 4, used at ec.rs:42-43
  Fp s = fp_sqrt_candidate(x);
  if (!fp_eq(fp_sqr(s), x)) return false;
  *r = s;
  return true;
}
struct GoodCurve {  // GoodCurve::Odd, ec.rs:34
  Fp a, b;
};
static bool good_curve_new_odd(const Fp& a, const Fp& bb, GoodCurve* c) {  // ec.rs:38-45
  Fp disc = fp_sub(fp_sqr(a), fp_add(fp_add(bb, bb), fp_add(bb, bb)));
  if (fp_is_zero(bb) || fp_is_zero(disc)) return false;
  Fp b, t;
  if (!fp_sqrt_host(bb, &b)) return false;
  if (!fp_sqrt_host(fp_add(fp_add(a, b), b), &t)) return false;
  c->a = a;
  c->b = b;
  return true;
}
static int two_adicity(Pt p, const GoodCurve& c) {  // utils.rs:356-365
  Fp a4 = fp_sqr(c.b);
  for (int i = 0; i < 2048; i++) {
    if (p.inf) return i;
    p = pt_add(p, p, c.a, a4);
  }
  return -1;
}
static Fp poly_eval_host(const std::vector<Fp>& c, const Fp& x) {
  Fp acc = fp_zero();
  for (size_t i = c.size(); i-- > 0;) acc = fp_add(fp_mul(acc, x), c[i]);
  return acc;
}

static void set_mempool_threshold(int device) {
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
    uint64_t thr = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
  }
}

static Tree* new_tree(size_t n, int parts, int device) {
  Tree* t = new Tree();
  t->device = device;
  t->log_n = ilog2(n);
  t->parts = parts;
  ECFFT_CUDA(cudaSetDevice(device));
  set_mempool_threshold(device);
  ECFFT_CUDA(cudaStreamCreateWithFlags(&t->stream, cudaStreamNonBlocking));
  t->f = t->dalloc(2 * n);
  ECFFT_CUDA(cudaMalloc((void**)&t->build_errors, sizeof(unsigned long long)));
  ECFFT_CUDA(cudaMemset(t->build_errors, 0, sizeof(unsigned long long)));
  return t;
}

// FFTree::new, src/fftree.rs:42-70: internal nodes layer by layer through the rational maps
static void fill_internal_nodes(Tree& t) {
  const size_t n = t.n();
  cudaStream_t st = t.stream;
  ECFFT_CUDA(cudaMemsetAsync(t.f, 0, sizeof(Fp), st));  // f[0] is the unused slot
  for (uint32_t k = 0; k < t.log_n; k++) {
    const RatMapHost& m = t.maps[k];
    Fp* coeff = nullptr;
    size_t cnt = m.num.size() + m.den.size();
    ECFFT_CUDA(cudaMallocAsync((void**)&coeff, (cnt ? cnt : 1) * sizeof(Fp), st));
    if (!m.num.empty()) ECFFT_CUDA(cudaMemcpyAsync(coeff, m.num.data(), m.num.size() * sizeof(Fp), cudaMemcpyHostToDevice, st));
    if (!m.den.empty()) ECFFT_CUDA(cudaMemcpyAsync(coeff + m.num.size(), m.den.data(), m.den.size() * sizeof(Fp), cudaMemcpyHostToDevice, st));
    k::ratmap_layer(t.f + (n >> (k + 1)), t.f + (n >> k), n >> (k + 1), coeff, (int)m.num.size(), coeff + m.num.size(), (int)m.den.size(), t.build_errors, st);
    ECFFT_CUDA(cudaStreamSynchronize(st));  // host coefficient vectors stay alive until the copy is done
    ECFFT_CUDA(cudaFreeAsync(coeff, st));
  }
}

Tree* tree_from_leaves(const Fp* leaves_dev_plain, size_t n, const std::vector<RatMapHost>& maps, int parts, int device) {
  if (!n || (n & (n - 1))) throw Error(ERR_NOT_POW2, "leaf count is not a power of two");
  if (ilog2(n) != maps.size()) throw Error(ERR_INVALID_ARG, "need log2(n) rational maps");
  Tree* t = new_tree(n, parts, device);
  try {
    t->maps = maps;
    ECFFT_CUDA(cudaMemcpyAsync(t->f + n, leaves_dev_plain, n * sizeof(Fp), cudaMemcpyDeviceToDevice, t->stream));
    fill_internal_nodes(*t);
    finish_tree(*t);
  } catch (...) {
    delete t;
    throw;
  }
  return t;
}

// Symmetric butterflies need every rational map in the form (x^2 + c1 x + c0)/x with c0 = beta^2 a
// nonzero square (the Good-curve 2-isogenies of src/ec.rs:84 are (x - b)^2/x): then the two nodes of
// a pair are s and beta^2/s.  Fills t.beta / t.sym_ok.
static void find_betas(Tree& t) {
  t.beta.clear();
  t.sym_ok = false;
  Fp two = fp_add(fp_one(), fp_one()), inv2 = fp_inv(two);
  for (const RatMapHost& m : t.maps) {
    if (m.den.size() != 2 || !fp_is_zero(m.den[0]) || !fp_eq(m.den[1], fp_one())) return;
    if (m.num.size() != 3 || !fp_eq(m.num[2], fp_one()) || fp_is_zero(m.num[0])) return;
    Fp beta = fp_mul(fp_neg(m.num[1]), inv2);
    if (!fp_eq(fp_sqr(beta), m.num[0])) {
      beta = fp_sqrt_candidate(m.num[0]);
      if (!fp_eq(fp_sqr(beta), m.num[0])) return;
    }
    t.beta.push_back(beta);
  }
  t.sym_ok = true;
}

// Normalised-butterfly tables of chain level k, derived from the top tree's f and the
// level's recombine matrices: symmetric form when the maps allow it, else the two-product form.
void build_norm_tables(Tree& t, uint32_t k) {
  if (k == 0) return;
  if (t.beta.size() != t.maps.size()) find_betas(t);
  const size_t n = t.n(), N = (size_t)1 << k, stride = n / N, hh = N / 2;
  cudaStream_t st = t.stream;
  Level& lv = t.levels[k];
  lv.sym = t.sym_ok && k::butterfly_mode() == 2;
  Fp* beta_dev = nullptr;
  Fp inv2L = fp_one();
  if (lv.sym) {
    // level with half-stride 2^j uses map k-2-j (the layer with 2^(j+2) nodes)
    std::vector<Fp> bj(k > 1 ? k - 1 : 1, fp_zero());
    Fp inv2 = fp_inv(fp_add(fp_one(), fp_one()));
    for (uint32_t j = 0; j + 1 < k; j++) {
      bj[j] = t.beta[k - 2 - j];
      inv2L = fp_mul(inv2L, inv2);
    }
    beta_dev = t.dalloc(bj.size());
    ECFFT_CUDA(cudaMemcpyAsync(beta_dev, bj.data(), bj.size() * sizeof(Fp), cudaMemcpyHostToDevice, st));
    ECFFT_CUDA(cudaStreamSynchronize(st));
  }
  const size_t per = lv.sym ? 1 : 2;
  for (int mu = 0; mu < 2; mu++) {
    lv.tw_r[mu] = t.dalloc(per * hh);
    lv.tw_d[mu] = t.dalloc(per * hh);
    lv.gam[mu] = t.dalloc(hh);
    lv.gami[mu] = t.dalloc(hh);
    if (lv.sym) {
      k::build_twiddles_sym(lv.tw_r[mu], lv.tw_d[mu], t.f, stride, hh, mu, beta_dev, t.build_errors, st);
      k::build_gamma_sym(lv.gam[mu], lv.rmat, t.f, stride, hh, mu, beta_dev, st);
    } else {
      k::build_twiddles(lv.tw_r[mu], lv.tw_d[mu], t.f, stride, hh, mu, st);
      k::build_gamma(lv.gam[mu], lv.rmat, hh, mu, st);
    }
    ECFFT_CUDA(cudaMemcpyAsync(lv.gami[mu], lv.gam[mu], hh * sizeof(Fp), cudaMemcpyDeviceToDevice, st));
    k::batch_inverse(lv.gami[mu], hh, st);
    if (lv.sym)
      k::mul_const(lv.gami[mu], lv.gami[mu], inv2L, hh, st);  // the decompose butterflies omit their 1/2
    else
      k::fold_sumform_prescale(lv.gami[mu], t.f, stride, hh, mu, st);
  }
  lv.gx = t.dalloc(hh);
  k::mul_strided(lv.gx, lv.gam[1], lv.xnn_s, 2, 1, hh, st);
  if (lv.sym && hh >= 2)  // centre constants: level 0 has one twiddle (entry 1) per moiety
    for (int mu = 0; mu < 2; mu++) {
      lv.ctr[mu] = t.dalloc(1);
      k::mul_strided(lv.ctr[mu], lv.tw_r[mu] + 1, lv.tw_d[1 - mu], 1, 1, 1, st);
    }
}

// One level of from_tree (src/fftree.rs:318-463): N = 2^k leaves = every (n/N)-th leaf of the top
// tree; its f layers are the same strided views of the top tree's layers (derive_subtree, :465-482).
static void build_level(Tree& t, uint32_t k) {
  const size_t n = t.n(), N = (size_t)1 << k, stride = n / N;
  cudaStream_t st = t.stream;
  Engine eng(t, st);
  Level& lv = t.levels[k];
  lv.log_n = k;

  Fp* s = eng.tmp(N);  // this level's leaves, contiguous
  k::copy_strided(s, t.f + n, N, stride, st);

  // <X^(N/2) on S> and inverse, fftree.rs:331-333
  lv.xnn_s = t.dalloc(N);
  lv.xnn_s_inv = t.dalloc(N);
  k::pow_u64(lv.xnn_s, s, N / 2, N, st);
  ECFFT_CUDA(cudaMemcpyAsync(lv.xnn_s_inv, lv.xnn_s, N * sizeof(Fp), cudaMemcpyDeviceToDevice, st));
  k::batch_inverse(lv.xnn_s_inv, N, st);

  // matrices, fftree.rs:341-363; entries never written stay identity
  lv.rmat = t.dalloc(4 * N);
  lv.dmat = t.dalloc(4 * N);
  {
    Fp id[4] = {fp_one(), fp_zero(), fp_zero(), fp_one()};
    // identity fill: two strided fills are enough (1 at slots 0 and 3, 0 at 1 and 2)
    ECFFT_CUDA(cudaMemsetAsync(lv.rmat, 0, 4 * N * sizeof(Fp), st));
    ECFFT_CUDA(cudaMemsetAsync(lv.dmat, 0, 4 * N * sizeof(Fp), st));
    std::vector<Fp> host(4 * N <= 64 ? 4 * N : 0);
    if (!host.empty()) {
      for (size_t i = 0; i < N; i++) memcpy(&host[4 * i], id, sizeof id);
      ECFFT_CUDA(cudaMemcpyAsync(lv.rmat, host.data(), host.size() * sizeof(Fp), cudaMemcpyHostToDevice, st));
      ECFFT_CUDA(cudaMemcpyAsync(lv.dmat, host.data(), host.size() * sizeof(Fp), cudaMemcpyHostToDevice, st));
      ECFFT_CUDA(cudaStreamSynchronize(st));
    } else {
      // only entries 0 and 1 keep the identity (layers with d >= 2 are overwritten below)
      ECFFT_CUDA(cudaMemcpyAsync(lv.rmat, id, sizeof id, cudaMemcpyHostToDevice, st));
      ECFFT_CUDA(cudaMemcpyAsync(lv.rmat + 4, id, sizeof id, cudaMemcpyHostToDevice, st));
      ECFFT_CUDA(cudaMemcpyAsync(lv.dmat, id, sizeof id, cudaMemcpyHostToDevice, st));
      ECFFT_CUDA(cudaMemcpyAsync(lv.dmat + 4, id, sizeof id, cudaMemcpyHostToDevice, st));
      ECFFT_CUDA(cudaStreamSynchronize(st));
    }
  }
  for (uint32_t kk = 0; kk < k; kk++) {
    const size_t d = N >> (kk + 1);
    if (d == 1) continue;  // fftree.rs:350-352
    const RatMapHost& m = t.maps[kk];
    Fp* den = nullptr;
    ECFFT_CUDA(cudaMallocAsync((void**)&den, (m.den.size() ? m.den.size() : 1) * sizeof(Fp), st));
    if (!m.den.empty()) ECFFT_CUDA(cudaMemcpyAsync(den, m.den.data(), m.den.size() * sizeof(Fp), cudaMemcpyHostToDevice, st));
    k::build_matrices(lv.rmat + 4 * d, lv.dmat + 4 * d, t.f + (n >> kk), stride, d, den, (int)m.den.size(), t.build_errors, st);
    ECFFT_CUDA(cudaStreamSynchronize(st));
    ECFFT_CUDA(cudaFreeAsync(den, st));
  }

  build_norm_tables(t, k);

  if (t.parts == PARTS_ENTER_ONLY || N == 1) {
    eng.release(s);
    return;
  }

  const size_t h = N / 2;
  lv.z0_s1 = t.dalloc(h);
  lv.z1_s0 = t.dalloc(h);
  lv.z0_inv_s1 = t.dalloc(h);
  lv.z1_inv_s0 = t.dalloc(h);
  lv.z0z0 = t.dalloc(N);
  lv.z1z1 = t.dalloc(N);
  Fp* s0 = eng.tmp(h);
  Fp* s1 = eng.tmp(h);
  k::deinterleave(s0, s1, s, h, st);

  if (N == 2) {  // base cases, fftree.rs:399-403, 454-458 (two leaves: a handful of host operations)
    Fp hs[2];
    ECFFT_CUDA(cudaMemcpyAsync(hs, s, 2 * sizeof(Fp), cudaMemcpyDeviceToHost, st));
    ECFFT_CUDA(cudaStreamSynchronize(st));
    Fp z0 = fp_sub(hs[1], hs[0]), z1 = fp_sub(hs[0], hs[1]);
    Fp z0i = fp_inv(z0), z1i = fp_inv(z1);
    Fp zz0[2] = {fp_sqr(hs[0]), fp_sqr(hs[0])}, zz1[2] = {fp_sqr(hs[1]), fp_sqr(hs[1])};
    ECFFT_CUDA(cudaMemcpyAsync(lv.z0_s1, &z0, sizeof(Fp), cudaMemcpyHostToDevice, st));
    ECFFT_CUDA(cudaMemcpyAsync(lv.z1_s0, &z1, sizeof(Fp), cudaMemcpyHostToDevice, st));
    ECFFT_CUDA(cudaMemcpyAsync(lv.z0_inv_s1, &z0i, sizeof(Fp), cudaMemcpyHostToDevice, st));
    ECFFT_CUDA(cudaMemcpyAsync(lv.z1_inv_s0, &z1i, sizeof(Fp), cudaMemcpyHostToDevice, st));
    ECFFT_CUDA(cudaMemcpyAsync(lv.z0z0, zz0, sizeof zz0, cudaMemcpyHostToDevice, st));
    ECFFT_CUDA(cudaMemcpyAsync(lv.z1z1, zz1, sizeof zz1, cudaMemcpyHostToDevice, st));
    ECFFT_CUDA(cudaStreamSynchronize(st));
    lv.has_z = true;
    eng.release(s); eng.release(s0); eng.release(s1);
    return;
  }

  const Level& sub = t.levels[k - 1];
  Fp* zero_h = eng.tmp(h / 2);
  ECFFT_CUDA(cudaMemsetAsync(zero_h, 0, (h / 2) * sizeof(Fp), st));
  // <Z_0 on S_1>: extend the subtree's vanishing polynomials, fftree.rs:386-393
  {
    Fp* a = eng.tmp(h);
    Fp* b = eng.tmp(h);
    k::interleave(a, zero_h, sub.z0_s1, h / 2, st);  // [0, y]
    k::interleave(b, sub.z1_s0, zero_h, h / 2, st);  // [y, 0]
    eng.extend(a, a, h, 1, S1);
    eng.extend(b, b, h, 1, S1);
    k::mul_bcast(lv.z0_s1, a, b, h, 1, st);
    eng.release(a);
    eng.release(b);
  }
  // <Z_1 on S_0> = vanish(S_1) at the even leaves, fftree.rs:395-397
  {
    Fp* z1_s = eng.tmp(N);
    eng.vanish(s1, z1_s, h, FORM_PLAIN);
    k::copy_strided(lv.z1_s0, z1_s, h, 2, st);
    eng.release(z1_s);
  }
  ECFFT_CUDA(cudaMemcpyAsync(lv.z0_inv_s1, lv.z0_s1, h * sizeof(Fp), cudaMemcpyDeviceToDevice, st));
  ECFFT_CUDA(cudaMemcpyAsync(lv.z1_inv_s0, lv.z1_s0, h * sizeof(Fp), cudaMemcpyDeviceToDevice, st));
  k::batch_inverse(lv.z0_inv_s1, h, st);
  k::batch_inverse(lv.z1_inv_s0, h, st);

  // <Z_0^2 mod X^(N/2) on S>, <Z_1^2 mod X^(N/2) on S>, fftree.rs:417-453
  {
    Fp* xnnnn_s = eng.tmp(N);
    Fp* xnnnn_s_inv = eng.tmp(N);
    k::pow_u64(xnnnn_s, s, N / 4, N, st);
    ECFFT_CUDA(cudaMemcpyAsync(xnnnn_s_inv, xnnnn_s, N * sizeof(Fp), cudaMemcpyDeviceToDevice, st));
    k::batch_inverse(xnnnn_s_inv, N, st);

    Fp* r_s0 = eng.tmp(h);
    Fp* r_s1 = eng.tmp(h);
    k::mul_bcast(r_s0, sub.z0z0, sub.z1z1, h, 1, st);                       // z0_rem_xnnnn_sq_s0
    eng.modular_reduce(r_s0, sub.xnn_s, nullptr, sub.z0z0, h, 1, r_s0);     // z0z0_rem_xnnnn_s0
    eng.extend(r_s0, r_s1, h, 1, S1);
    Fp* rem = eng.tmp(N);                                                  // z0z0_rem_xnnnn_s
    k::interleave(rem, r_s0, r_s1, h, st);
    Fp* q = eng.tmp(N);
    k::sqr_sub_mul(q, lv.z0_s1, 1, lv.xnn_s, rem, xnnnn_s_inv, N, st);      // ((Z0 - X^(N/2))^2 - rem) / X^(N/4)
    eng.modular_reduce(q, xnnnn_s, nullptr, rem, N, 1, q);
    k::muladd(lv.z0z0, rem, xnnnn_s, q, N, st);                            // rem + X^(N/4) * q
    k::sqr_sub_mul(q, lv.z1_s0, 0, lv.xnn_s, nullptr, nullptr, N, st);      // (Z1 - X^(N/2))^2
    eng.modular_reduce(q, lv.xnn_s, nullptr, lv.z0z0, N, 1, lv.z1z1);
    eng.release(xnnnn_s); eng.release(xnnnn_s_inv); eng.release(r_s0); eng.release(r_s1);
    eng.release(rem); eng.release(q);
  }
  lv.has_z = true;
  eng.release(zero_h); eng.release(s); eng.release(s0); eng.release(s1);
}

void finish_tree(Tree& t) {
  const size_t n = t.n();
  t.levels.assign(t.log_n + 1, Level());
  // leaves of the 2-leaf chain level (VANISH base case)
  t.base_leaf0 = fp_zero();
  t.base_leaf1 = fp_zero();
  if (t.log_n >= 1) {
    ECFFT_CUDA(cudaMemcpyAsync(&t.base_leaf0, t.f + n, sizeof(Fp), cudaMemcpyDeviceToHost, t.stream));
    ECFFT_CUDA(cudaMemcpyAsync(&t.base_leaf1, t.f + n + n / 2, sizeof(Fp), cudaMemcpyDeviceToHost, t.stream));
    ECFFT_CUDA(cudaStreamSynchronize(t.stream));
  }
  for (uint32_t k = 0; k <= t.log_n; k++) build_level(t, k);
  // the reference panics on these inputs (`unwrap()` at src/fftree.rs:57-58, :361); here they are an error code
  unsigned long long bad = 0;
  if (t.build_errors) ECFFT_CUDA(cudaMemcpyAsync(&bad, t.build_errors, sizeof bad, cudaMemcpyDeviceToHost, t.stream));
  ECFFT_CUDA(cudaStreamSynchronize(t.stream));
  if (bad) throw Error(ERR_INVALID_ARG, "leaves / rational maps are degenerate: " + std::to_string(bad) + " zero denominators, singular matrices or nodes at a fixed point");
}

// Fp::build_fftree, reference src/lib.rs:39-85 (constants :45-59 in hex; tests check them against
// the reference's decimal literals).
Tree* build_secp256k1(size_t n, int parts, int device) {
  if (!n || (n & (n - 1))) throw Error(ERR_NOT_POW2, "n is not a power of two");
  const uint32_t log_n = ilog2(n);
  const uint32_t gen_two_adicity = 36;
  if (log_n >= gen_two_adicity) throw Error(ERR_TOO_LARGE, "FFTree size is too large for the generator (log2 n >= 36)");

  GoodCurve curve;
  if (!good_curve_new_odd(fp_from_hex("44eae664a07c69e1c7d7821cacf2a3ccca446568bd32b2a48166309c5c4297e5"),
                          fp_from_hex("649cd342698de65c9bc86f1ece3beb99197d6715a53bdb5609cc937a16154ca8"), &curve))
    throw Error(ERR_INVALID_ARG, "bad curve constants");
  Pt offset, gen;
  offset.inf = gen.inf = false;
  offset.x = fp_from_hex("e9850041b13ea03fadc4bee2afd2959604bf64c290bf3fc15165f15163fd5431");
  offset.y = fp_from_hex("110b996c1374482d0a6b9055a21dc8af9a098495b902b3663322f53ee416d65f");
  gen.x = fp_from_hex("5b4b3e43cd5d95fba244389bb8655539cf8d527f331697e2e93ea60ef50ad5c4");
  gen.y = fp_from_hex("a30fcedca51e68850478e0905816b86d88b79d7b549f4a340016e31de71ded06");
  const Fp a4 = fp_sqr(curve.b);
  for (uint32_t i = 0; i < gen_two_adicity - log_n; i++) gen = pt_add(gen, gen, curve.a, a4);  // lib.rs:67-70

  // find_isogeny_chain, ec.rs:177-189, keeping only each isogeny's x-map r
  std::vector<RatMapHost> maps;
  {
    Pt g = gen;
    GoodCurve c = curve;
    int kk = two_adicity(g, c);
    if (kk != (int)log_n) throw Error(ERR_INVALID_ARG, "generator does not have order n");
    for (int i = 0; i < kk; i++) {
      // good_isogeny, Odd branch, ec.rs:75-88
      Fp bb = fp_sqr(c.b), b2 = fp_add(c.b, c.b), b4 = fp_add(b2, b2);
      Fp a_prime = fp_add(fp_add(c.a, b4), b2);
      Fp ab = fp_mul(c.a, c.b), ab4 = fp_add(fp_add(ab, ab), fp_add(ab, ab));
      Fp bb2 = fp_add(bb, bb), bb8 = fp_add(fp_add(bb2, bb2), fp_add(bb2, bb2));
      GoodCurve cod;
      if (!good_curve_new_odd(a_prime, fp_add(ab4, bb8), &cod)) throw Error(ERR_INVALID_ARG, "isogeny codomain is not a good curve");
      RatMapHost r, hmap;
      r.num = {bb, fp_neg(b2), fp_one()};
      r.den = {fp_zero(), fp_one()};
      hmap.num = {fp_neg(bb), fp_zero(), fp_one()};
      hmap.den = {fp_zero(), fp_zero(), fp_one()};
      // Isogeny::map, ec.rs:344-358 (g = 0)
      Pt gp;
      Fp rd = poly_eval_host(r.den, g.x), hd = poly_eval_host(hmap.den, g.x);
      if (fp_is_zero(rd) || fp_is_zero(hd)) {
        gp = pt_infinity();
      } else {
        gp.inf = false;
        gp.x = fp_mul(poly_eval_host(r.num, g.x), fp_inv(rd));
        gp.y = fp_mul(fp_mul(poly_eval_host(hmap.num, g.x), fp_inv(hd)), g.y);
      }
      if (two_adicity(g, c) != two_adicity(gp, cod) + 1) throw Error(ERR_INVALID_ARG, "isogeny does not halve the generator's order");
      maps.push_back(r);
      g = gp;
      c = cod;
    }
  }

  Tree* t = new_tree(n, parts, device);
  try {
    t->maps = maps;
    cudaStream_t st = t->stream;
    // table of 2^j * G for the leaf kernel
    std::vector<Fp> gtab(2 * (log_n ? log_n : 1));
    Pt gj = gen;
    for (uint32_t j = 0; j < log_n; j++) {
      gtab[2 * j] = gj.x;
      gtab[2 * j + 1] = gj.y;
      gj = pt_add(gj, gj, curve.a, a4);
    }
    Fp* dtab = nullptr;
    ECFFT_CUDA(cudaMallocAsync((void**)&dtab, gtab.size() * sizeof(Fp), st));
    ECFFT_CUDA(cudaMemcpyAsync(dtab, gtab.data(), gtab.size() * sizeof(Fp), cudaMemcpyHostToDevice, st));
    k::build_leaves(t->f + n, n, curve.a, a4, offset.x, offset.y, dtab, log_n, st);  // lib.rs:72-78
    ECFFT_CUDA(cudaStreamSynchronize(st));
    ECFFT_CUDA(cudaFreeAsync(dtab, st));
    fill_internal_nodes(*t);
    finish_tree(*t);
  } catch (...) {
    delete t;
    throw;
  }
  return t;
}

}  // namespace ecfft
