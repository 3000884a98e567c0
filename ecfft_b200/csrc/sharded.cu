// Per-rank schedule of the fully sharded multi-GPU ENTER over peer-mapped memory.
//
// ENTER's recursion (reference src/fftree.rs:143-161) splits the coefficient vector into contiguous halves,
// so with G ranks the bottom log2(n/G) depths of chunk g are an independent ENTER(n/G) on rank g.  For the
// top log2(G) depths a rank holds a contiguous chunk of c = n/G elements of one length-h vector (r = h/c
// ranks per vector).  Butterfly levels with half-stride >= c pair this rank's chunk with ONE other rank's
// chunk: the cross-level kernel loads the partner's operands straight from the partner's arena over NVLink
// and computes this rank's output of every pair (twiddles depend only on the position modulo the
// half-stride).  Levels with half-stride < c run the tile kernel on the chunk.  The combine
// (src/fftree.rs:155-159) reads u0, u1 from the u-vector's rank and v0, v1 from the v-vector's rank.
// Every buffer a call produces has its own arena slot and all ranks number slots and synchronisation steps
// alike, so a peer's buffer of the same step sits at the same offset of its arena; ordering is by
// stream-ordered u64 flags (k::mg_sync).  A call starts by waiting until every peer has finished the previous
// call (its reads of this rank's arena included) and ends by publishing that itself has — a device-side
// barrier.  No host round trip, no send/recv, no collective.
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
size_t peer_arena_bytes(size_t n, int world) {
  if (world <= 0 || (world & (world - 1)) || n == 0 || (n & (n - 1)) || n % (size_t)world) throw Error(ERR_NOT_POW2, "n and world size must be powers of two");
  return MG_FLAG_BYTES + peer_slots(world) * (n / (size_t)world) * sizeof(Fp);
}

void enter_peer(const Engine& eng, const Fp* chunk, size_t n, int rank, int world, void* const* bases,
                unsigned long long epoch, Fp* out_chunk) {
  if (world <= 0 || (world & (world - 1)) || n == 0 || (n & (n - 1)) || n % (size_t)world) throw Error(ERR_NOT_POW2, "n and world size must be powers of two");
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
  const unsigned timeout_ms = peer_timeout_ms();
  size_t next_slot = 0, sid = 0;
  auto slot = [&](int r, size_t idx, size_t off = 0) { return (Fp*)((char*)bases[r] + MG_FLAG_BYTES) + idx * c + off; };
  auto flag = [&](int r, size_t s) { return (unsigned long long*)bases[r] + s; };
  auto sync = [&](int a, int b) {  // publish everything enqueued so far, then wait for the same step of ranks a, b
    if (sid >= MG_DONE_FLAG) throw Error(ERR_INVALID_ARG, "peer arena: too many synchronisation steps");
    k::mg_sync(flag(rank, sid), epoch, a != rank ? flag(a, sid) : nullptr, (b != rank && b != a) ? flag(b, sid) : nullptr, timeout_ms, st);
    sid++;
  };

  // nobody may still be reading this rank's arena from the previous call when it gets overwritten
  if (epoch > 1) k::mg_wait_all(bases, world, rank, MG_DONE_FLAG, epoch - 1, timeout_ms, st);
  size_t sA = next_slot++;
  eng.enter_range(chunk, slot(rank, sA), c, 1, c);
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
      sW = next_slot++;
      k::mul_bcast(slot(rank, sW), slot(rank, sA), lv.gami[0] + pos0, c, 1, st);
    }
    for (uint32_t j = log_h; j-- > log_c;) {  // decompose levels whose pairs straddle two ranks
      const int bit = (k >> (j - log_c)) & 1, peer = rank ^ (1 << (j - log_c));
      sync(peer, peer);
      const size_t sN = next_slot++;
      k::mg_cross(lv, 0, j, bit, pos0 - ((size_t)bit << j), slot(rank, sW), slot(peer, sW), c, slot(rank, sN), st);
      sW = sN;
    }
    {
      const size_t sN = next_slot++;
      k::extend_sub(lv, slot(rank, sW), slot(rank, sN), log_c, st, local_pre);
      sW = sN;
    }
    for (uint32_t j = log_c; j < log_h; j++) {  // recombine levels that straddle two ranks
      const int bit = (k >> (j - log_c)) & 1, peer = rank ^ (1 << (j - log_c));
      sync(peer, peer);
      const size_t sN = next_slot++;
      k::mg_cross(lv, 1, j, bit, pos0 - ((size_t)bit << j), slot(rank, sW), slot(peer, sW), c, slot(rank, sN), st);
      sW = sN;
    }
    // ---- combine: output rank kp of the block takes i in [kp c/2, (kp+1) c/2) from the u- and the v-rank
    const size_t half = c / 2;
    const int kp = rank - block0;
    const int usrc = block0 + kp / 2, vsrc = block0 + r + kp / 2;
    const size_t off = (size_t)(kp % 2) * half;
    sync(usrc, vsrc);
    const bool last = 2 * m > n;
    Fp* dst = out_chunk;
    size_t sNext = 0;
    if (!last) {
      sNext = next_slot++;
      dst = slot(rank, sNext);
    }
    k::mg_combine(lv, (size_t)kp * half, slot(usrc, sA, off), slot(vsrc, sA, off), slot(usrc, sW, off), slot(vsrc, sW, off), half, dst, st);
    sA = sNext;
  }
  k::mg_sync(flag(rank, MG_DONE_FLAG), epoch, nullptr, nullptr, 0, st);  // this rank reads no peer memory any more
}

}  // namespace ecfft
