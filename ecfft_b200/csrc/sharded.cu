// Per-rank schedules of the fully sharded multi-GPU ENTER and EXIT over peer-mapped memory.
//
// ENTER's recursion (reference src/fftree.rs:143-161) splits the coefficient vector into contiguous halves,
// so with G ranks the bottom log2(n/G) depths of chunk g are an independent ENTER(n/G) on rank g.  For the
// top log2(G) depths a rank holds a contiguous chunk of c = n/G elements of one length-h vector (r = h/c
// ranks per vector).  Butterfly levels with half-stride >= c pair this rank's chunk with ONE other rank's
// chunk: the cross-level kernel loads the partner's operands straight from the partner's arena over NVLink
// and computes this rank's output of every pair (twiddles depend only on the position modulo the
// half-stride).  Levels with half-stride < c run the tile kernel on the chunk.  The combine
// (src/fftree.rs:155-159) reads u0, u1 from the u-vector's rank and v0, v1 from the v-vector's rank.
//
// EXIT (src/fftree.rs:200-224) is the mirror image: its top log2(G) depths run MOD (two REDCs = four EXTENDs
// of the half-length vectors of even / odd samples, src/fftree.rs:232-259, 277-281) on vectors spread over the
// ranks with the same cross-level kernels, then split (u0 | v0) and hand each half to half of the ranks; below,
// every rank runs an independent EXIT(n/G).
//
// Every buffer a call produces for its peers has its own arena slot and all ranks number slots and
// synchronisation steps alike, so a peer's buffer of the same step sits at the same offset of its arena;
// ordering is by stream-ordered u64 flags (k::mg_sync).  A call starts by waiting until every peer has finished
// the previous call (its reads of this rank's arena included) and ends by publishing that itself has — a
// device-side barrier.  No host round trip, no send/recv, no collective.
#include <cstdlib>

#include "engine.h"

namespace ecfft {

static inline uint32_t ilog2(size_t n) {
  uint32_t l = 0;
  while (n >>= 1) l++;
  return l;
}

// How long a flag wait may last before the waiting kernel traps (a CUDA error on the next call rather than a
// hung GPU): ECFFT_B200_PEER_TIMEOUT_MS, default 20000, 0 = wait for ever (ranks that may reach a call far
// apart in time — a tree build or data loading on one of them — should raise it or use 0).
unsigned peer_timeout_ms() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ECFFT_B200_PEER_TIMEOUT_MS");
    long t = e ? atol(e) : 20000;
    v = t < 0 ? 20000 : (t > 0x7fffffff ? 0x7fffffff : (int)t);
  }
  return (unsigned)v;
}

static size_t peer_slots(int world) {
  // A0; per top depth with r ranks per vector: W_pre, one per straddling level (2 log2 r), W_local, A_next
  size_t slots = 1;
  for (int r = 1; r < world; r *= 2) slots += 2 + 2 * ilog2((size_t)r) + 1;
  return slots;
}
static void check_shape(size_t n, int world) {
  if (world <= 0 || (world & (world - 1)) || n == 0 || (n & (n - 1)) || n % (size_t)world) throw Error(ERR_NOT_POW2, "n and world size must be powers of two");
}
size_t peer_arena_bytes(size_t n, int world) {
  check_shape(n, world);
  return MG_FLAG_BYTES + peer_slots(world) * (n / (size_t)world) * sizeof(Fp);
}
static size_t peer_exit_slots(int world) {
  // per top depth with r ranks per vector (L = log2 r): two REDCs of two EXTENDs (input + 2L cross levels + local
  // each), u0 and v0
  size_t slots = 0;
  for (int r = world; r > 1; r /= 2) slots += 2 * 2 * (2 + 2 * ilog2((size_t)r)) + 2;
  return slots;
}
static size_t peer_exit_slot_elems(size_t n, int world) { return n / (size_t)world / 2; }
size_t peer_exit_arena_bytes(size_t n, int world) {
  check_shape(n, world);
  return MG_FLAG_BYTES + peer_exit_slots(world) * peer_exit_slot_elems(n, world) * sizeof(Fp);
}

namespace {
// bookkeeping of one call: arena slots and synchronisation steps, numbered alike on all ranks
struct PeerCtx {
  cudaStream_t st;
  int rank, world;
  void* const* bases;
  unsigned long long epoch;
  size_t slot_elems;
  unsigned timeout_ms;
  size_t next_slot = 0, sid = 0;
  Fp* slot(int r, size_t idx, size_t off = 0) const { return (Fp*)((char*)bases[r] + MG_FLAG_BYTES) + idx * slot_elems + off; }
  unsigned long long* flag(int r, size_t s) const { return (unsigned long long*)bases[r] + s; }
  size_t new_slot() { return next_slot++; }
  // publish everything enqueued so far, then wait for the same step of ranks a, b
  void sync(int a, int b) {
    if (sid >= MG_STATUS_FLAG) throw Error(ERR_INVALID_ARG, "peer arena: too many synchronisation steps");
    k::mg_sync(flag(rank, sid), epoch, a != rank ? flag(a, sid) : nullptr, (b != rank && b != a) ? flag(b, sid) : nullptr, timeout_ms, st,
               flag(rank, MG_STATUS_FLAG), (unsigned)sid);
    sid++;
  }
  void begin() {  // nobody may still be reading this rank's arena from the previous call when it gets overwritten
    if (epoch > 1) k::mg_wait_all(bases, world, rank, MG_DONE_FLAG, epoch - 1, timeout_ms, st, flag(rank, MG_STATUS_FLAG));
  }
  void end() { k::mg_sync(flag(rank, MG_DONE_FLAG), epoch, nullptr, nullptr, 0, st); }  // this rank reads no peer memory any more
};

// All butterfly levels of the EXTEND (source -> target moiety) of a vector of 2^log_h elements held in chunks
// of 2^log_len elements by 2^(log_h - log_len) consecutive ranks, this rank being number k of them: the
// decompose levels whose pairs straddle two ranks, the rank-local levels (tile kernel, optional pre-scale
// riding its first stage), the straddling recombine levels.  sW: slot of this rank's (pre-scaled) chunk;
// returns the slot of the result, which still lacks the Gamma^target scaling.
size_t extend_levels_peer(PeerCtx& c, const Level& lv, uint32_t log_h, uint32_t log_len, int k, size_t sW, Moiety source, Moiety target,
                          const Fp* local_pre) {
  const size_t len = (size_t)1 << log_len, pos0 = (size_t)k << log_len;
  for (uint32_t j = log_h; j-- > log_len;) {  // decompose levels whose pairs straddle two ranks
    const int bit = (k >> (j - log_len)) & 1, peer = c.rank ^ (1 << (j - log_len));
    c.sync(peer, peer);
    const size_t sN = c.new_slot();
    k::mg_cross(lv, 0, j, bit, pos0 - ((size_t)bit << j), c.slot(c.rank, sW), c.slot(peer, sW), len, c.slot(c.rank, sN), c.st, source, target);
    sW = sN;
  }
  {
    const size_t sN = c.new_slot();
    k::extend_sub(lv, c.slot(c.rank, sW), c.slot(c.rank, sN), log_len, c.st, local_pre, source, target);
    sW = sN;
  }
  for (uint32_t j = log_len; j < log_h; j++) {  // recombine levels that straddle two ranks
    const int bit = (k >> (j - log_len)) & 1, peer = c.rank ^ (1 << (j - log_len));
    c.sync(peer, peer);
    const size_t sN = c.new_slot();
    k::mg_cross(lv, 1, j, bit, pos0 - ((size_t)bit << j), c.slot(c.rank, sW), c.slot(peer, sW), len, c.slot(c.rank, sN), c.st, source, target);
    sW = sN;
  }
  return sW;
}

// CUDA loads a kernel's code on its first launch, and that load can wait for the whole context to go idle — which
// never happens while another rank's stream OF THE SAME PROCESS (virtual ranks on one GPU: the tests) spins on a
// flag this thread has yet to publish.  So every kernel of the schedules runs once, on scratch data, before the
// first flag wait is enqueued (once per device; with one process per GPU it is a millisecond of set-up).
void warm_up(const Engine& eng, bool with_exit) {
  static PerDeviceOnce once_enter, once_exit;
  (with_exit ? once_exit : once_enter).run([&] {
    cudaStream_t st = eng.st;
    const size_t nw = eng.t.n() < 2048 ? eng.t.n() : 2048;
    if (nw < 8) return;
    Fp* s = eng.tmp(2 * nw + 64);
    ECFFT_CUDA(cudaMemsetAsync(s, 0, (2 * nw + 64) * sizeof(Fp), st));
    Fp *a = s, *b = s + nw, *f = s + 2 * nw;   // f: 64 elements of flags / tiny operands
    const Level& lv = eng.level_for(8);
    if (!lv.has_norm()) throw Error(ERR_MISSING_TABLES, "sharded schedules need the normalised butterfly tables");
    eng.enter_range(a, b, nw, 1, nw);
    for (int phase = 0; phase < 2; phase++) k::mg_cross(lv, phase, 1, 0, 0, f, f + 2, 2, f + 4, st, S0, S1);
    k::extend_sub(lv, f, f + 8, 1, st, lv.gami[0], S0, S1);
    k::extend_sub(lv, f, f + 8, 2, st, nullptr, S1, S0);
    k::mg_combine(lv, 0, f, f + 2, f + 4, f + 6, 2, f + 8, st);
    k::mul_bcast(f + 8, f, lv.gami[0], 2, 1, st);
    void* bases1[1] = {f + 16};
    k::mg_wait_all(bases1, 1, 0, 0, 1, 1, st);
    k::mg_sync((unsigned long long*)(f + 16), 1, nullptr, nullptr, 0, st);
    if (with_exit) {
      eng.exit(a, b, nw);
      k::mul_strided(f + 8, lv.gami[0], f, 2, 0, 2, st);
      k::dot2_strided(f + 8, f, 2, lv.gami[0], f + 4, lv.gami[1], 2, st);
      k::sub_mul_strided(f + 8, f, 2, f + 4, lv.gami[0], 2, st);
      ECFFT_CUDA(cudaMemcpyAsync(f + 12, f, 2 * sizeof(Fp), cudaMemcpyDefault, st));
    }
    ECFFT_CUDA(cudaStreamSynchronize(st));
    eng.release(s);
  });
}
}  // namespace

void enter_peer(const Engine& eng, const Fp* chunk, size_t n, int rank, int world, void* const* bases,
                unsigned long long epoch, Fp* out_chunk) {
  check_shape(n, world);
  if (rank < 0 || rank >= world) throw Error(ERR_INVALID_ARG, "bad rank");
  const size_t c = n / (size_t)world;
  if (world > 1 && c < 2) throw Error(ERR_INVALID_ARG, "sharded ENTER needs at least 2 coefficients per rank");
  const uint32_t log_c = ilog2(c);
  cudaStream_t st = eng.st;
  eng.level_for(n);
  if (world == 1) {
    eng.enter_range(chunk, out_chunk, n, 1, n);
    return;
  }
  warm_up(eng, false);
  PeerCtx ctx{st, rank, world, bases, epoch, c, peer_timeout_ms()};
  ctx.begin();
  size_t sA = ctx.new_slot();
  // Streams for the rank-local ENTER, measured on 2 / 4 / 8 B200s at n = 2^22 (profiles/r02_y_*, r02_ag_*): chunks of 2^20
  // elements and more run best on two streams (7.73 vs 8.08 ms on one at 2 GPUs, 4.59 vs 4.72 / 4.68 ms on one / four at 4),
  // chunks of 2^19 on the automatic four (8 GPUs: 3.12 ms, 3.13 on two, 3.27 on one).  (Before programmatic dependent
  // launch was restricted by grid size the order at 8 GPUs was the reverse — 3.18 / 3.44 / 3.50 ms — for the reason given
  // at pdl_mode in sym_kernel.cu.)
  eng.enter_range(chunk, ctx.slot(rank, sA), c, 1, c, c >= ((size_t)1 << 20) ? 2 : 0);
  int r = 1;
  for (size_t m = 2 * c; m <= n; m *= 2, r *= 2) {
    const Level& lv = eng.level_for(m);
    if (!lv.has_norm()) throw Error(ERR_MISSING_TABLES, "sharded ENTER needs the normalised butterfly tables");
    const size_t h = m / 2;
    const uint32_t log_h = ilog2(h);
    const int k = rank % r;
    const size_t pos0 = (size_t)k * c;
    const int block0 = (rank / r / 2) * 2 * r;
    // ---- EXTEND -> S1 of the vector this rank holds a chunk of
    size_t sW;
    const Fp* local_pre = nullptr;
    if (r == 1) {          // the whole vector is local: the pre-scale rides the tile kernel's first stage
      sW = sA;
      local_pre = lv.gami[0];
    } else {
      sW = ctx.new_slot();
      k::mul_bcast(ctx.slot(rank, sW), ctx.slot(rank, sA), lv.gami[0] + pos0, c, 1, st);
    }
    sW = extend_levels_peer(ctx, lv, log_h, log_c, k, sW, S0, S1, local_pre);
    // ---- combine: output rank kp of the block takes i in [kp c/2, (kp+1) c/2) from the u- and the v-rank
    const size_t half = c / 2;
    const int kp = rank - block0;
    const int usrc = block0 + kp / 2, vsrc = block0 + r + kp / 2;
    const size_t off = (size_t)(kp % 2) * half;
    ctx.sync(usrc, vsrc);
    const bool last = 2 * m > n;
    Fp* dst = out_chunk;
    size_t sNext = 0;
    if (!last) {
      sNext = ctx.new_slot();
      dst = ctx.slot(rank, sNext);
    }
    k::mg_combine(lv, (size_t)kp * half, ctx.slot(usrc, sA, off), ctx.slot(vsrc, sA, off), ctx.slot(usrc, sW, off), ctx.slot(vsrc, sW, off), half, dst, st);
    sA = sNext;
  }
  ctx.end();
}

void exit_peer(const Engine& eng, const Fp* chunk, size_t n, int rank, int world, void* const* bases,
               unsigned long long epoch, Fp* out_chunk) {
  check_shape(n, world);
  if (rank < 0 || rank >= world) throw Error(ERR_INVALID_ARG, "bad rank");
  const size_t c = n / (size_t)world;
  cudaStream_t st = eng.st;
  eng.level_for(n);
  if (world == 1) {
    eng.exit(chunk, out_chunk, n);
    return;
  }
  if (c < 4) throw Error(ERR_INVALID_ARG, "sharded EXIT needs at least 4 evaluations per rank");
  const size_t cc = c / 2;              // samples of one moiety in a chunk
  const uint32_t log_cc = ilog2(cc);
  // every level's EXIT tables exist before anything that waits for a peer is enqueued (building them synchronises
  // the tree's stream with the host)
  for (size_t m = n; m >= 4; m /= 2) {
    const Level& lv = eng.level_for(m);
    if (!lv.z0z0) throw Error(ERR_MISSING_TABLES, "exit: tree was built without the Z tables");
    if (!eng.exit_tabs(lv)) throw Error(ERR_MISSING_TABLES, "sharded EXIT needs the symmetric butterfly tables");
  }
  warm_up(eng, true);
  PeerCtx ctx{st, rank, world, bases, epoch, peer_exit_slot_elems(n, world), peer_timeout_ms()};
  ctx.begin();
  // rank-local scratch (never read by a peer)
  Fp* next_buf[2] = {eng.tmp(c), eng.tmp(c)};
  Fp* hb_even = eng.tmp(cc);
  Fp* hb_odd = eng.tmp(cc);
  Fp* h1 = eng.tmp(cc);
  const Fp* cur = chunk;
  int nb = 0;
  int r = world;
  for (size_t m = n; r > 1; m /= 2, r /= 2) {
    const Level& lv = eng.level_for(m);
    if (!lv.z0z0) throw Error(ERR_MISSING_TABLES, "exit: tree was built without the Z tables");
    if (!eng.exit_tabs(lv)) throw Error(ERR_MISSING_TABLES, "sharded EXIT needs the symmetric butterfly tables");
    const uint32_t log_h = ilog2(m / 2);
    const int k = rank % r, group0 = rank - k;
    const size_t pos0 = (size_t)k * cc;     // position of this rank's samples in the half-length vectors
    size_t sU = 0;
    // MOD = REDC, x z0z0, REDC with a = xnn_s (fftree.rs:206-210, 277-281); the tables are the level's prebuilt
    // {a0inv * gami (* c_even), -(gam * a_odd * zinv), zinv (* c_odd)} of the fused form (engine.cu)
    for (int ri = 0; ri < 2; ri++) {
      const Fp* even = ri == 0 ? cur : hb_even;
      const Fp* odd = ri == 0 ? cur + 1 : hb_odd;
      const size_t stride = ri == 0 ? 2 : 1;
      Fp* const* tab = lv.exit_tab[ri];
      // t0 = e0 / a0 (already carrying the pre-scale), g1 = EXTEND(t0 -> S1)
      size_t sW = ctx.new_slot();
      k::mul_strided(ctx.slot(rank, sW), tab[0] + pos0, even, stride, 0, cc, st);
      sW = extend_levels_peer(ctx, lv, log_h, log_cc, k, sW, S0, S1, nullptr);
      // h1 = (e1 - g1 a1) / Z0 (fftree.rs:253-255), as one two-product dot
      Fp* h1_dst = ri == 0 ? hb_odd : h1;
      k::dot2_strided(h1_dst, odd, stride, tab[2] + pos0, ctx.slot(rank, sW), tab[1] + pos0, cc, st);
      // h0 = EXTEND(h1 -> S0)
      size_t sV = ctx.new_slot();
      k::mul_bcast(ctx.slot(rank, sV), h1_dst, lv.gami[S1] + pos0, cc, 1, st);
      sV = extend_levels_peer(ctx, lv, log_h, log_cc, k, sV, S1, S0, nullptr);
      if (ri == 0) {
        k::mul_bcast(hb_even, ctx.slot(rank, sV), lv.gam[S0] + pos0, cc, 1, st);
      } else {  // M[::2] = u0 (fftree.rs:206-210): this rank's piece of it, where its new owner will fetch it
        sU = ctx.new_slot();
        k::mul_bcast(ctx.slot(rank, sU), ctx.slot(rank, sV), lv.gam[S0] + pos0, cc, 1, st);
      }
    }
    // v0 = (e0 - u0) / x^(m/2) on S0 (fftree.rs:215-219)
    const size_t sV0 = ctx.new_slot();
    k::sub_mul_strided(ctx.slot(rank, sV0), cur, 2, ctx.slot(rank, sU), lv.exit_a0inv + pos0, cc, st);
    // u0 goes to the first half of the vector's ranks, v0 to the second half: rank k' of a half takes its two
    // pieces from ranks 2k', 2k'+1 of the parent group
    const int r2 = r / 2, newk = k % r2;
    const bool second = k >= r2;
    const int srcA = group0 + 2 * newk, srcB = srcA + 1;
    const size_t part = second ? sV0 : sU;
    ctx.sync(srcA, srcB);
    Fp* nx = next_buf[nb];
    nb ^= 1;
    ECFFT_CUDA(cudaMemcpyAsync(nx, ctx.slot(srcA, part), cc * sizeof(Fp), cudaMemcpyDefault, st));
    ECFFT_CUDA(cudaMemcpyAsync(nx + cc, ctx.slot(srcB, part), cc * sizeof(Fp), cudaMemcpyDefault, st));
    cur = nx;
  }
  // the peers' pieces have been fetched: nothing below touches peer memory
  ctx.end();
  eng.exit(cur, out_chunk, c);
  eng.release(next_buf[0]);
  eng.release(next_buf[1]);
  eng.release(hb_even);
  eng.release(hb_odd);
  eng.release(h1);
}

}  // namespace ecfft
