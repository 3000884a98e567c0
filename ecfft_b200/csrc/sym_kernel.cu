// sm_100a EXTEND kernel for the symmetric (one-product) butterflies — the dominant kernel of the
// engine — with ENTER's combine (reference src/fftree.rs:155-159) fused into its last pass.
// See DESIGN.md 4.1.
//
// Butterfly network of EXTEND (flattening of extend_impl, src/fftree.rs:72-120): log2(h) decompose
// levels (half-strides h/2 .. 1) then log2(h) recombine levels (1 .. h/2), in place, natural order.
// With g = (s - beta)/(s + beta) the two nodes of a pair carry +g and -g, so a recombine butterfly is
// y_p = x_p + g x_q, y_q = x_p - g x_q and a decompose butterfly x_p = y_p + y_q, x_q = (y_p - y_q)/g
// (halvings and the diagonal Gamma scalings folded into one pre- and one post-scale per element).
//
// A CTA owns a tile of 2^log_t elements closed under a group of consecutive levels, kept in shared
// memory (low and high 16 bytes of the elements in separate arrays: conflict-free 16-byte accesses).
// A thread takes FOUR elements into registers and runs two levels on them (four at the centre of the
// inner pass: decompose 1, 0 then recombine 0, 1) per shared-memory round trip.  The tile arrives by
// cp.async (no registers held across the HBM latency); the pre-scale is applied by the first stage as
// it reads the tile; after the last stage an epilogue reads the tile back in natural order and either
// stores it (post-scale, canonical reduction, coalesced) or, in ENTER, performs the combine
//   out[2i] = u0[i] + v0[i] xnn[2i],  out[2i+1] = gam[i] u1[i] + gx[i] v1[i],
// for which a tile holds the same positions of the two sibling vectors u and v.  The first load and the
// last store can be stride-2 views and the store can add a second operand (e1 * Z + x * post), which is
// how REDC's de-interleave, pointwise steps and interleave (src/fftree.rs:232-259) ride the two EXTENDs.
#include <cuda.h>   // CUtensorMap and the cuTensorMapEncodeTiled prototype only: the entry point is looked up at run time
#include <cstdlib>

#include "engine.h"

namespace ecfft {
namespace k {

enum : uint32_t { OP_D_HI = 1, OP_D_LO = 2, OP_R_LO = 4, OP_R_HI = 8, OP_C_LO = 16 };

struct TileSoA {
  uint4* s;
  uint32_t T;
  __device__ __forceinline__ Fp ld(uint32_t e) const {
    uint4 a = s[e], b = s[T + e];
    Fp r;
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    return r;
  }
  __device__ __forceinline__ void st(uint32_t e, const Fp& x) const {
    s[e] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
    s[T + e] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
  }
};

__device__ __forceinline__ void sym_d_pair(Fp& a, Fp& b, const Fp& ginv) {  // x_p = y_p + y_q, x_q = (y_p - y_q)/g
  Fp d = fp_sub_lazy2_f(a, b);
  a = fp_add_lazy_f(a, b);
  b = fp_mul_lazy(ginv, d);
}
__device__ __forceinline__ void sym_r_pair(Fp& a, Fp& b, const Fp& g) {     // y_p = x_p + g x_q, y_q = x_p - g x_q
  Fp t = fp_mul_lazy(g, b);
  b = fp_sub_lazy2_f(a, t);
  a = fp_add_lazy_f(a, t);
}
// centre of the network: decompose then recombine at the same level on the same pair is
// [[1, 1], [1, -1]] diag(1, g_target/g_source) [[1, 1], [1, -1]]: one product with c = g_target/g_source
__device__ __forceinline__ void sym_c_pair(Fp& a, Fp& b, const Fp& c) {
  Fp t = fp_mul_lazy(c, fp_sub_lazy2_f(a, b));
  Fp sum = fp_add_lazy_f(a, b);
  a = fp_add_lazy_f(sum, t);
  b = fp_sub_lazy2_f(sum, t);
}
__device__ __forceinline__ void cp_async16(void* smem, const void* gmem) {
  unsigned sa = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gmem) : "memory");
}

// ---- TMA tile load (ECFFT_B200_TMA=1): the tile arrives as bulk tensor copies issued by ONE thread and
// completes on an mbarrier, instead of 16 cp.async per thread.  The buffer is seen as a 3-D tensor of u32:
// [row = element >> row_shift][column = element & (2^row_shift - 1)][8 limbs]; a box of {4 limbs, <= 256 columns,
// 2^krows rows} lands densely in shared memory, which is exactly one half (low or high 16 bytes) of the split
// tile layout.  Elements beyond the batch are zero-filled by the copy engine.
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@!p bra WAIT_%=;\n\t}"
      ::"r"((unsigned)__cvta_generic_to_shared(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* smem, const CUtensorMap* map, int c0, int c1, int c2, unsigned long long* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
               ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(map), "r"(c0), "r"(c1), "r"(c2),
                 "r"((unsigned)__cvta_generic_to_shared(bar)) : "memory");
}
__device__ __forceinline__ void bulk_prefetch_l2(const void* p, unsigned bytes) {   // one instruction per contiguous range, no registers held
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}
__device__ __forceinline__ void prefetch_l1(const void* p) { asm volatile("prefetch.global.L1 [%0];" ::"l"(p)); }
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

// tile element -> (offset from the tile's first global element, position within its vector)
struct TileMap {
  uint32_t log_c, cmask, rmask, krows, row_shift, log_h, pos0;
  __device__ __forceinline__ void map(uint32_t e, unsigned long long& goff, uint32_t& pos) const {
    const uint32_t rr = e >> log_c, c = e & cmask;
    const uint32_t rel = ((rr & rmask) << row_shift) + c;
    pos = pos0 + rel;
    goff = ((unsigned long long)(rr >> krows) << log_h) + rel;
  }
};

// One stage of a tile: the quad of item q differs in tile-index bits b_lo (level jl) and b_hi (level jh); ops
// selects which of the level operations run.  When both levels are used jl = jh - 1, so the two low-level pairs
// share a twiddle.  Single-level stages (odd level counts) pair bit b_hi only; b_lo is any other bit.
struct Stage {
  uint32_t ops, jh, jl, b1, b2, S_lo, S_hi, mh, ml;
};
struct StagePlan {
  uint32_t nD, mid, j_base, odd, npairs, nstages;
  __device__ __forceinline__ StagePlan(const SymParams& p) {
    const uint32_t nlev = p.lvl_hi - p.lvl_lo;
    mid = (p.do_d && p.do_r && nlev >= 2) ? 1u : 0u;   // levels lvl_lo+1, lvl_lo run D, D+R (one product), R in one stage
    j_base = p.lvl_lo + (mid ? 2 : 0);                 // lowest level of the plain D / R stages
    const uint32_t cnt = p.lvl_hi - j_base;
    odd = cnt & 1;
    npairs = cnt >> 1;
    nD = p.do_d ? odd + npairs : 0;
    const uint32_t nR = p.do_r ? odd + npairs : 0;
    nstages = nD + mid + nR;
  }
  __device__ __forceinline__ Stage stage(const SymParams& p, uint32_t sidx) const {
    Stage g;
    if (sidx < nD) {
      if (odd && sidx == 0) { g.ops = OP_D_HI; g.jh = p.lvl_hi - 1; g.jl = 0; }
      else { const uint32_t u = sidx - odd; g.ops = OP_D_HI | OP_D_LO; g.jh = p.lvl_hi - 1 - odd - 2 * u; g.jl = g.jh - 1; }
    } else if (mid && sidx == nD) {
      g.ops = OP_D_HI | OP_C_LO | OP_R_HI; g.jh = p.lvl_lo + 1; g.jl = p.lvl_lo;
    } else {
      const uint32_t t = sidx - nD - mid;
      if (t < npairs) { g.ops = OP_R_LO | OP_R_HI; g.jl = j_base + 2 * t; g.jh = g.jl + 1; }
      else { g.ops = OP_R_HI; g.jh = p.lvl_hi - 1; g.jl = 0; }
    }
    const bool two = (g.ops & (OP_D_LO | OP_R_LO | OP_C_LO)) != 0;
    const uint32_t b_hi = g.jh + p.boff;
    const uint32_t b_lo = two ? g.jl + p.boff : (b_hi == 0 ? 1u : b_hi - 1);
    g.b1 = b_lo < b_hi ? b_lo : b_hi;
    g.b2 = b_lo < b_hi ? b_hi : b_lo;
    g.S_lo = 1u << b_lo;
    g.S_hi = 1u << b_hi;
    g.mh = (1u << g.jh) - 1;
    g.ml = (1u << g.jl) - 1;
    return g;
  }
};
__device__ __forceinline__ uint32_t quad_e0(uint32_t q, const Stage& g) {
  uint32_t e0 = ((q >> g.b1) << (g.b1 + 1)) | (q & ((1u << g.b1) - 1));   // zero bit at b1
  return ((e0 >> g.b2) << (g.b2 + 1)) | (e0 & ((1u << g.b2) - 1));        // and at b2
}

// Twiddle prefetch (ECFFT_B200_TW_PREFETCH = 1: into L1, 2: into L2): while a stage computes, the lines the NEXT
// stage's butterflies of this thread will load through the read-only path are requested, so the per-butterfly
// twiddle loads of the strided passes (one 32-byte value per butterfly, no reuse inside a tile) stop
// exposing the L2 / DRAM latency.  Levels with small tables (< 256 twiddles) stay cached anyway and are skipped.
template <int NT>
__device__ __forceinline__ void prefetch_stage(const SymParams& p, const TileMap& tm, const Stage& g, uint32_t T, bool first) {
  const uint32_t hmask = (1u << p.log_h) - 1;
  const bool hi_big = g.jh >= 8, lo_big = g.jl >= 8;
  if (!hi_big && !lo_big && !first) return;
#pragma unroll 1
  for (uint32_t q = threadIdx.x; q < T / 4; q += NT) {
    const uint32_t e0 = quad_e0(q, g), e1 = e0 + g.S_lo;
    unsigned long long g0;
    uint32_t pa, pb;
    tm.map(e0, g0, pa);
    tm.map(e1, g0, pb);
    const Fp *t0 = nullptr, *t1 = nullptr, *t2 = nullptr;
    if (hi_big) {
      const Fp* tab = (g.ops & OP_D_HI) && !(g.ops & OP_R_HI) ? p.tw_d : p.tw_r;   // the centre stage loads both: its levels are tiny
      t0 = tab + (1u << g.jh) + (pa & g.mh);
      t1 = tab + (1u << g.jh) + (pb & g.mh);
    }
    if (lo_big && (g.ops & (OP_D_LO | OP_R_LO))) t2 = ((g.ops & OP_D_LO) ? p.tw_d : p.tw_r) + (1u << g.jl) + (pa & g.ml);
    if (p.pf == 1) {
      if (t0) { prefetch_l1(t0); prefetch_l1(t1); }
      if (t2) prefetch_l1(t2);
    } else {
      if (t0) { prefetch_l2(t0); prefetch_l2(t1); }
      if (t2) prefetch_l2(t2);
    }
    if (first) {   // the pre-scale table of the first stage: positions of the whole quad
      uint32_t pc, pd;
      tm.map(e0 + g.S_hi, g0, pc);
      tm.map(e1 + g.S_hi, g0, pd);
      prefetch_l1(p.pre + (pa & hmask)); prefetch_l1(p.pre + (pb & hmask));
      prefetch_l1(p.pre + (pc & hmask)); prefetch_l1(p.pre + (pd & hmask));
    }
  }
}

// Bulk L2 prefetch of what a STRIDED tile will read through the read-only path later (ECFFT_B200_L2PF, one thread,
// one cp.async.bulk.prefetch.L2 per contiguous range): the twiddles of its levels — level j needs, for every
// combination of the row bits below j, 2^log_c consecutive entries — and, for the pair tile of ENTER's last pass,
// the rows of u0 / v0 and of the combine tables.  In those passes a butterfly's twiddle has no reuse inside the tile
// and the top recursion depths have few vectors to share lines with, so without this the loads go to DRAM when the
// stage needs them (ncu source view: 23 % of the warps' time in such a launch waited on them).  Measured slower
// all the same (see launch_sym): kept as a knob with its numbers, off by default.
__device__ __forceinline__ void prefetch_tile_operands(const SymParams& p, uint32_t pos0, unsigned long long gbase) {
  // the 32 lanes of one warp share the ranges (range number mod 32 == lane)
  const uint32_t lane = threadIdx.x & 31;
  uint32_t cnt = 0;
  auto pf = [&](const void* ptr, uint32_t bytes) {
    if ((cnt++ & 31) == lane) bulk_prefetch_l2(ptr, bytes);
  };
  const uint32_t run = (uint32_t)sizeof(Fp) << p.log_c;
  const uint32_t cpos = pos0 & ((1u << p.row_shift) - 1);          // column offset of the tile inside a row
  for (uint32_t j = p.lvl_lo; j < p.lvl_hi; j++) {
    const uint32_t nr = 1u << (j - p.row_shift);
    for (uint32_t r = 0; r < nr; r++) {
      const uint32_t idx = (1u << j) + (r << p.row_shift) + cpos;
      if (p.do_d) pf(p.tw_d + idx, run);
      if (p.do_r) pf(p.tw_r + idx, run);
    }
  }
  if (p.comb) {
    const uint32_t hmask = (1u << p.log_h) - 1;
    for (uint32_t r = 0; r < (1u << p.krows); r++) {
      const unsigned long long g = gbase + ((unsigned long long)r << p.row_shift);
      const uint32_t i = (pos0 + (r << p.row_shift)) & hmask;
      pf(p.A + g, run);
      pf(p.A + g + ((unsigned long long)1 << p.log_h), run);
      pf(p.gam + i, run);
      pf(p.gx + i, run);
      if (p.ce0) {
        pf(p.ce0 + i, run);
        pf(p.ce1 + i, run);
      } else {
        pf(p.xnn + 2 * i, 2 * run);
      }
    }
  } else if (p.post && !p.E) {
    const uint32_t hmask = (1u << p.log_h) - 1;
    for (uint32_t r = 0; r < (1u << p.krows); r++) pf(p.post + ((pos0 + (r << p.row_shift)) & hmask), run);
  }
}

// One tile of one pass.  Packed tiles start at element gbase_packed; strided tiles are tile `tile` of vector
// (pair) w.  FLOW: the data buffers may have been written by other CTAs of the SAME launch, so every read of
// them goes to L2 (cp.async.cg, ld.global.cg), never through L1.  TMA: the tile load is a bulk tensor copy.
template <int NT, bool FLOW, bool TMA>
__device__ __forceinline__ void sym_tile(const SymParams& p, uint4* smem_raw, unsigned long long w, unsigned long long tile,
                                         unsigned long long gbase_packed, const CUtensorMap* tmap = nullptr,
                                         unsigned long long* mbar = nullptr) {
  const uint32_t T = 1u << p.log_t;
  const TileSoA s{smem_raw, T};
  const uint32_t hmask = (1u << p.log_h) - 1;
  TileMap tm;
  unsigned long long gbase;
  if (p.packed) {
    tm = TileMap{p.log_t, T - 1, 0u, 0u, 0u, p.log_h, 0u};
    gbase = gbase_packed;
  } else {
    const uint32_t ncg_log = p.row_shift - p.log_c;
    const uint32_t cg = (uint32_t)(tile & ((1ull << ncg_log) - 1));
    const uint32_t q_hi = (uint32_t)(tile >> ncg_log);
    const uint32_t pos0 = (q_hi << p.lvl_hi) + (cg << p.log_c);  // position within the vector of tile element 0
    tm = TileMap{p.log_c, (1u << p.log_c) - 1, (1u << p.krows) - 1, p.krows, p.row_shift, p.log_h, pos0};
    gbase = ((p.pair ? 2 * w : w) << p.log_h) + pos0;
  }
  const StagePlan plan(p);
  Stage cur = plan.stage(p, 0);
  // l2pf = 1: issued while the tile load is in flight; 2: after the tile has landed (the copy engine serves both)
  if (!FLOW && p.l2pf == 1 && !p.packed && (threadIdx.x >> 5) == 1) prefetch_tile_operands(p, tm.pos0, gbase);

  // ---- tile load
  if (TMA) {
    // one thread: arm the barrier with the tile's bytes, issue the boxes (low halves, then high halves)
    if (threadIdx.x == 0) {
      const uint32_t rs = p.packed ? p.log_t : p.row_shift;           // the tensor map's row length is 2^rs elements
      const uint32_t lc = p.packed ? p.log_t : p.log_c, kr = p.packed ? 0u : p.krows;
      const uint32_t bc_log = lc < 8 ? lc : 8;                        // box: 2^bc_log columns x 2^krb rows (dense in shared memory,
      const uint32_t krb = lc > 8 ? 0u : kr;                          //  so a row wider than a box goes row by row)
      const uint32_t nchunk = 1u << (lc - bc_log), nrow = 1u << (kr - krb), npair = p.pair ? 2u : 1u;
      const int c1 = (int)(gbase & ((1ull << rs) - 1)), c2 = (int)(gbase >> rs);
      const int prow = p.pair ? (int)(1u << (p.log_h - rs)) : 0;     // rows from the u-vector's tile to the v-vector's
      if (p.tma_fence) asm volatile("fence.proxy.async;" ::: "memory");   // debugging aid (ECFFT_B200_TMA=2)
      mbar_expect_tx(mbar, T * (uint32_t)sizeof(Fp));
      for (uint32_t hf = 0; hf < 2; hf++)
        for (uint32_t pb = 0; pb < npair; pb++)
          for (uint32_t rw = 0; rw < nrow; rw++)
            for (uint32_t cc = 0; cc < nchunk; cc++)
              tma_load_3d(&s.s[hf * T + (pb << (kr + lc)) + (rw << lc) + (cc << bc_log)], tmap, (int)(hf * 4), c1 + (int)(cc << bc_log),
                          c2 + (int)pb * prow + (int)rw, mbar);
    }
    if (p.pf) prefetch_stage<NT>(p, tm, cur, T, p.pre != nullptr);
    mbar_wait(mbar, 0);
  } else {
    // cp.async straight into the split shared-memory layout
    for (uint32_t e = threadIdx.x; e < T; e += NT) {
      unsigned long long goff;
      uint32_t pos;
      tm.map(e, goff, pos);
      const unsigned long long g = gbase + goff;
      if (g < p.total) {
        const uint4* src = reinterpret_cast<const uint4*>(p.in + ((g << p.in_shift) + p.in_off));
        cp_async16(&s.s[e], src);
        cp_async16(&s.s[T + e], src + 1);
      } else {
        s.st(e, fp_zero());
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    if (p.pf) prefetch_stage<NT>(p, tm, cur, T, p.pre != nullptr);
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
  }

  if (!FLOW && p.l2pf == 2 && !p.packed && (threadIdx.x >> 5) == 1) prefetch_tile_operands(p, tm.pos0, gbase);

  // ---- stages
  for (uint32_t sidx = 0; sidx < plan.nstages; sidx++) {
    const Stage g = cur;
    if (sidx + 1 < plan.nstages) {
      cur = plan.stage(p, sidx + 1);
      if (p.pf) prefetch_stage<NT>(p, tm, cur, T, false);
    }
    const uint32_t ops = g.ops, S_lo = g.S_lo, S_hi = g.S_hi, mh = g.mh, ml = g.ml;
    const bool first = sidx == 0 && p.pre != nullptr;
    const Fp* d_hi = p.tw_d + (1u << g.jh);
    const Fp* d_lo = p.tw_d + (1u << g.jl);
    const Fp* r_hi = p.tw_r + (1u << g.jh);
    const Fp* r_lo = p.tw_r + (1u << g.jl);
#pragma unroll 1
    for (uint32_t q = threadIdx.x; q < T / 4; q += NT) {
      const uint32_t e0 = quad_e0(q, g);
      const uint32_t e1 = e0 + S_lo, e2 = e0 + S_hi, e3 = e1 + S_hi;
      unsigned long long g0, g1, g2, g3;
      uint32_t pa, pb, pc, pd;
      tm.map(e0, g0, pa);
      tm.map(e1, g1, pb);
      tm.map(e2, g2, pc);
      tm.map(e3, g3, pd);
      Fp x0 = s.ld(e0), x1 = s.ld(e1), x2 = s.ld(e2), x3 = s.ld(e3);
      if (first) {
        x0 = fp_mul_lazy(x0, fp_load_ro(p.pre + (pa & hmask)));
        x1 = fp_mul_lazy(x1, fp_load_ro(p.pre + (pb & hmask)));
        x2 = fp_mul_lazy(x2, fp_load_ro(p.pre + (pc & hmask)));
        x3 = fp_mul_lazy(x3, fp_load_ro(p.pre + (pd & hmask)));
      }
      if (ops & OP_D_HI) {
        sym_d_pair(x0, x2, fp_load_ro(d_hi + (pa & mh)));
        sym_d_pair(x1, x3, fp_load_ro(d_hi + (pb & mh)));
      }
      if (ops & OP_D_LO) {
        const Fp gi = fp_load_ro(d_lo + (pa & ml));
        sym_d_pair(x0, x1, gi);
        sym_d_pair(x2, x3, gi);
      }
      if (ops & OP_C_LO) {
        const Fp c_lo = fp_load_ro(p.ctr);
        sym_c_pair(x0, x1, c_lo);
        sym_c_pair(x2, x3, c_lo);
      }
      if (ops & OP_R_LO) {
        const Fp gg = fp_load_ro(r_lo + (pa & ml));
        sym_r_pair(x0, x1, gg);
        sym_r_pair(x2, x3, gg);
      }
      if (ops & OP_R_HI) {
        sym_r_pair(x0, x2, fp_load_ro(r_hi + (pa & mh)));
        sym_r_pair(x1, x3, fp_load_ro(r_hi + (pb & mh)));
      }
      s.st(e0, x0); s.st(e1, x1); s.st(e2, x2); s.st(e3, x3);
    }
    __syncthreads();
  }

  if (!p.comb) {
    // store: post-scale (or REDC's h1 = e1 * Z + x * post, one reduction), canonical form, coalesced
#pragma unroll 1
    for (uint32_t e = threadIdx.x; e < T; e += NT) {
      unsigned long long g;
      uint32_t pos;
      tm.map(e, g, pos);
      g += gbase;
      if (g >= p.total) continue;
      const uint32_t i = pos & hmask;
      Fp x = s.ld(e);
      if (p.split) {
        // EXIT's split fused into the depth's last EXTEND (src/fftree.rs:206-220): u0 | (e0 - u0) / x^(n/2)
        const Fp u0 = fp_canon(fp_mul_lazy(x, fp_load_ro(p.post + i)));
        const Fp e0 = FLOW ? fp_load_cg(p.E + ((g << p.e_shift) + p.e_off)) : fp_load(p.E + ((g << p.e_shift) + p.e_off));
        Fp* o = p.out + ((g >> p.log_h) << (p.log_h + 1)) + i;
        fp_store(o, u0);
        fp_store(o + ((size_t)1 << p.log_h), fp_mul(fp_sub(e0, u0), fp_load_ro(p.Z + i)));
        continue;
      }
      if (p.E)
        x = fp_dot2_lazy(FLOW ? fp_load_cg(p.E + ((g << p.e_shift) + p.e_off)) : fp_load(p.E + ((g << p.e_shift) + p.e_off)), fp_load_ro(p.Z + i), x, fp_load_ro(p.post + i));
      else if (p.Z && p.post)   // MEXTEND inside VANISH (src/fftree.rs:133-134, 304): Z[i] + x * post[i], one reduction
        x = fp_muladd_lazy(fp_load_ro(p.Z + i), x, fp_load_ro(p.post + i));
      else if (p.post)
        x = fp_mul_lazy(x, fp_load_ro(p.post + i));
      fp_store(p.out + ((g << p.out_shift) + p.out_off), fp_canon(x));
    }
  } else {
    // ENTER combine (src/fftree.rs:155-159).  The tile holds the unscaled EXTEND of u at element eu and of
    // v at ev for the same position i; u0, v0 come back from global memory (this CTA's own input
    // when the whole EXTEND ran in this launch, else the depth's input vector).
    const uint32_t ush = p.packed ? p.log_h : p.log_t - 1;
#pragma unroll 1
    for (uint32_t idx = threadIdx.x; idx < T / 2; idx += NT) {
      const uint32_t eu = ((idx >> ush) << (ush + 1)) | (idx & ((1u << ush) - 1));
      const uint32_t ev = eu + (1u << ush);
      unsigned long long gu, gv;
      uint32_t pu, pv;
      tm.map(eu, gu, pu);
      tm.map(ev, gv, pv);
      gu += gbase;
      gv += gbase;
      if (gv >= p.total) continue;
      const uint32_t i = pu & hmask;
      Fp* o = p.out + (gu - i) + 2ull * i;
      const Fp u0 = FLOW ? fp_load_cg(p.A + gu) : fp_load(p.A + gu), v0 = FLOW ? fp_load_cg(p.A + gv) : fp_load(p.A + gv);
      if (p.ce0)   // folded forms: the scale the next depth's EXTEND wants (and / or the one this depth's input carries) in the tables
        fp_store(o, fp_canon(fp_dot2_lazy(fp_load_ro(p.ce0 + i), u0, fp_load_ro(p.ce1 + i), v0)));
      else
        fp_store(o, fp_canon(fp_muladd_lazy(u0, v0, fp_load_ro(p.xnn + 2 * i))));
      const Fp u1 = s.ld(eu), v1 = s.ld(ev);
      fp_store(o + 1, fp_canon(fp_dot2_lazy(fp_load_ro(p.gam + i), u1, fp_load_ro(p.gx + i), v1)));
    }
  }
}

template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_extend_sym(const __grid_constant__ SymParams p) {
  extern __shared__ __align__(128) uint4 smem_raw[];
  // Programmatic dependent launch (ECFFT_B200_PDL=1): the next pass's CTAs may be scheduled while this grid
  // drains — everything below still waits for the previous grid's results (griddepcontrol.wait returns once
  // the prerequisite grid has completed and its memory is visible).  No-ops without the launch attribute.
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (p.cond && __ldcg(p.cond) == 0) return;
  if (p.packed)
    sym_tile<NT, false, false>(p, smem_raw, 0, 0, (unsigned long long)blockIdx.x << p.log_t);
  else
    sym_tile<NT, false, false>(p, smem_raw, blockIdx.x % p.nv, blockIdx.x / p.nv, 0);
}
// The same pass with the tile load done by the TMA unit (tensor map encoded per launch by launch_sym)
template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_extend_sym_tma(const __grid_constant__ SymParams p, const __grid_constant__ CUtensorMap tmap) {
  extern __shared__ __align__(128) uint4 smem_raw[];
  __shared__ __align__(8) unsigned long long mbar;
  if (threadIdx.x == 0) mbar_init(&mbar, 1);
  __syncthreads();
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
  if (p.cond && __ldcg(p.cond) == 0) return;
  if (p.packed)
    sym_tile<NT, false, true>(p, smem_raw, 0, 0, (unsigned long long)blockIdx.x << p.log_t, &tmap, &mbar);
  else
    sym_tile<NT, false, true>(p, smem_raw, blockIdx.x % p.nv, blockIdx.x / p.nv, 0, &tmap, &mbar);
}

// ------------------------------------------------------------------------------------------------------
// Flow kernel: every pass of an ENTER (all recursion depths: EXTEND passes, fused and plain combines) in ONE
// launch.  Persistent CTAs (one per resident slot) take tiles from a global queue; the queue lists the passes
// in order and, inside a pass, the tiles in the order their inputs become ready.  A tile waits until the
// aligned block of the previous pass's output it reads is complete (a counter per block, release/acquire at
// GPU scope) — butterfly levels and the combine only ever read inside such a block, so no kernel boundary
// and no grid-wide barrier is needed, and the tail of one pass overlaps the head of the next.
// Deadlock freedom: a tile only waits for tiles with smaller queue numbers, which were taken earlier by CTAs
// that are running (only running CTAs take tiles).
// ------------------------------------------------------------------------------------------------------
static constexpr int FLOW_MAX_PASSES = 112;
struct FlowDesc {
  SymParams pass[FLOW_MAX_PASSES];
  uint32_t npass, total_tiles;
  unsigned int* queue;      // next tile number
  unsigned int* counters;   // per (pass, block) completed-tile counts
  unsigned long long* stats;  // null, or 4 diagnostic accumulators (ecfft_flow_stats)
};
static_assert(sizeof(FlowDesc) <= 32764, "flow descriptor must fit the kernel parameter space");

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu(unsigned* p, unsigned v) {
  asm volatile("fence.acq_rel.gpu;\n\tred.relaxed.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int NT, int MINB>
__global__ void __launch_bounds__(NT, MINB) k_sym_flow(const __grid_constant__ FlowDesc d) {
  extern __shared__ uint4 smem_raw[];
  __shared__ unsigned s_next[2];
  __shared__ unsigned* s_prev_sig;                     // counter of the tile this CTA finished last (signalled late)
  __shared__ long long s_acc[4];                       // diagnostics: wait, body, signal cycles, tiles
  uint32_t k = 0, it = 0;
  if (threadIdx.x == 0) {
    s_next[0] = atomicAdd(d.queue, 1u);
    s_acc[0] = s_acc[1] = s_acc[3] = 0;
  }
  if (threadIdx.x == NT - 1) {
    s_prev_sig = nullptr;
    s_acc[2] = 0;
  }
  for (;; it++) {
    __syncthreads();                                   // s_next[it & 1] is there; every store of the previous tile has been issued
    const unsigned g = s_next[it & 1];
    if (threadIdx.x == NT - 1 && s_prev_sig) {         // the previous tile's results are complete: publish (release, GPU scope)
      const long long t0 = d.stats ? clock64() : 0;
      red_release_gpu(s_prev_sig, 1u);
      if (d.stats) s_acc[2] += clock64() - t0;
    }
    if (g >= d.total_tiles) break;
    if (threadIdx.x == 0) s_next[(it + 1) & 1] = atomicAdd(d.queue, 1u);   // the next tile's number arrives while this one runs
    while (g >= d.pass[k].tile_begin + d.pass[k].ntiles) k++;
    const SymParams& p = d.pass[k];
    // queue position inside the pass -> (block b, vector w, tile i of the block) -> tile t_v of vector w
    const uint32_t sq = g - p.tile_begin;
    const uint32_t i = sq & ((1u << p.ord_tpb_log) - 1), t1 = sq >> p.ord_tpb_log;
    const uint32_t w = t1 % p.ord_nv, b = t1 / p.ord_nv;
    const unsigned long long t_v = ((unsigned long long)b << p.ord_tpb_log) | i;
    unsigned long long first_in, first_out, gbase_packed = 0;
    if (p.kind == 1 || p.packed) {
      gbase_packed = ((unsigned long long)w << p.ord_log_v) + (t_v << p.log_t);
      first_in = first_out = gbase_packed;
    } else {
      const uint32_t ncg_log = p.row_shift - p.log_c;
      const unsigned long long pos0 = ((t_v >> ncg_log) << p.lvl_hi) + ((t_v & ((1ull << ncg_log) - 1)) << p.log_c);
      first_in = ((unsigned long long)(p.pair ? 2 * w : w) << p.log_h) + pos0;
      first_out = p.pair ? ((unsigned long long)w << (p.log_h + 1)) + 2 * pos0 : first_in;
    }
    if (threadIdx.x == NT - 1) s_prev_sig = d.counters + p.sig_base + (unsigned)(first_out >> p.sig_shift);
    if (p.dep_need) {
      if (threadIdx.x == 0) {
        const long long t0 = d.stats ? clock64() : 0;
        const unsigned* c = d.counters + p.dep_base + (unsigned)(first_in >> p.dep_shift);
        if (ld_acquire_gpu(c) < p.dep_need) {
          // a dependency that never completes would be a bug in the plan: trap (a CUDA error) rather than hang the GPU
          unsigned long long ta, tb;
          asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(ta));
          unsigned ns = 64;
          while (ld_acquire_gpu(c) < p.dep_need) {
            __nanosleep(ns);                            // back off: many waiting CTAs poll the same few counters
            if (ns < 1024) ns *= 2;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tb));
            if (tb - ta > 10000000000ull) __trap();
          }
        }
        if (d.stats) s_acc[0] += clock64() - t0;
      }
      __syncthreads();
    }
    if (d.stats && threadIdx.x == 0) s_acc[1] -= clock64();
    if (p.kind == 0) {
      sym_tile<NT, true, false>(p, smem_raw, w, t_v, gbase_packed);
    } else {
      // combine-only pass (src/fftree.rs:155-159): outputs [first_out, first_out + T) of the batch
      const uint32_t T = 1u << p.log_t;
      const unsigned long long pair0 = gbase_packed >> 1, hm = (1ull << p.log_h) - 1;
#pragma unroll 1
      for (uint32_t e = threadIdx.x; e < T / 2; e += NT) {
        const unsigned long long idx = pair0 + e, blk = idx >> p.log_h, ii = idx & hm, off = blk << (p.log_h + 1);
        if (off + (1ull << p.log_h) + ii >= p.total) continue;
        const Fp u0 = fp_load_cg(p.A + off + ii), v0 = fp_load_cg(p.A + off + (1ull << p.log_h) + ii);
        fp_store(p.out + off + 2 * ii, fp_canon(fp_muladd_lazy(u0, v0, fp_load_ro(p.xnn + 2 * ii))));
        const Fp u1 = fp_load_cg(p.in + off + ii), v1 = fp_load_cg(p.in + off + (1ull << p.log_h) + ii);
        fp_store(p.out + off + 2 * ii + 1, fp_canon(fp_dot2_lazy(fp_load_ro(p.gam + ii), u1, fp_load_ro(p.gx + ii), v1)));
      }
    }
    if (d.stats && threadIdx.x == 0) { s_acc[1] += clock64(); s_acc[3]++; }
  }
  if (d.stats) {   // diagnostics (ecfft_flow_stats): cycles thread 0 spent waiting for inputs / in tile bodies, signal cycles
    if (threadIdx.x == 0) {
      atomicAdd(d.stats + 0, (unsigned long long)s_acc[0]);
      atomicAdd(d.stats + 1, (unsigned long long)s_acc[1]);
      atomicAdd(d.stats + 3, (unsigned long long)s_acc[3]);
    }
    if (threadIdx.x == NT - 1) atomicAdd(d.stats + 2, (unsigned long long)s_acc[2]);
  }
}

// Launch shapes (ECFFT_B200_SYM_VARIANT): 0 = 128 threads, 5 CTAs/SM (96 registers), 1024-element tile;
// 1 = 128 threads, 4 CTAs/SM (126 registers; default); 2 = 256 threads, 3 CTAs/SM; 3 = 256 threads, 2 CTAs/SM, 2048-element tile;
// 4 = 64 threads, 10 CTAs/SM, 512-element tile (shorter CTAs: smaller end-of-launch drain, one level less per pass).
static int sym_variant() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ECFFT_B200_SYM_VARIANT");
    // default 1 (126 registers, 4 CTAs/SM): measured 0.5 % (n = 2^22) to 5 % (2^20) faster than shape 0 on B200
    // (profiles/r02_e_*, r02_f_*): two products interleave without register pressure
    v = e ? atoi(e) : 1;
    if (v < 0 || v > 4) v = 1;
  }
  return v;
}
// ECFFT_B200_SYM_AUTO (default 1): launches of 2048 tiles and more take shape 0 (five CTAs of 96 registers per SM), smaller
// ones the default shape 1 — measured after the aligned reduction freed registers (profiles/r02_ak_ab_auto_shape.txt,
// r02_p_ab_shape_final.txt): ENTER 2^22 13.84 -> 13.69 ms, EXIT 2^22 30.02 -> 29.96 ms, while at 2^19 shape 0 loses
// (1.73 -> 1.80 ms), so the grid decides.
static bool sym_auto_shape() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ECFFT_B200_SYM_AUTO");
    v = e ? (atoi(e) != 0) : 1;
  }
  return v != 0;
}
static uint32_t sym_log_tile() { return sym_variant() == 3 ? 11 : sym_variant() == 4 ? 9 : 10; }
uint32_t flow_log_tile() { return sym_log_tile(); }
bool flow_enabled() {
  static int v = -1;
  if (v < 0) {
    // Off by default: measured on B200 (profiles/r02_c_*, r02_d_*) the flow launch is 4-10 % slower than one
    // launch per pass at every size — its tile body runs from a dynamically indexed pass table (no
    // constant-bank operands: 96 registers with spills, or 126 registers at 4 CTAs/SM) and the per-tile
    // queue / counter traffic costs ~2.5 %, while the drains it removes are already filled by the
    // two-stream ENTER.
    const char* e = getenv("ECFFT_B200_FLOW");
    v = e ? (atoi(e) != 0) : 0;
  }
  return v != 0;
}

// Programmatic dependent launch, by grid size (ECFFT_B200_PDL: 0 = never, 1 = this rule (default), 2 = always).
// Measured per size on one box (profiles/r02_ae_ab_pdl_tma_streams.txt, r02_ad_*): launches of up to ~64 tiles gain
// (ENTER 2^16 0.660 -> 0.613 ms, EXIT 2^16 2.81 -> 2.62 ms, ENTER -> EXIT 2^12 8 %); launches of 128-1024 tiles LOSE,
// badly where the grid is a fraction of one wave (EXIT 2^18 3.91 -> 6.72 ms, EXIT 2^19 5.35 -> 7.72 ms, ENTER 2^18
// 1.13 -> 1.58 ms, ENTER 2^19 on four streams 1.77 -> 2.10 ms): the dependent grid's CTAs are placed while the running
// grid still holds its slots, so an under-filled grid ends up packed on few SMs instead of spread over all 148;
// multi-wave launches are neutral to +1 % (ENTER 2^22 14.19 -> 14.17 ms, EXIT 2^22 30.95 -> 30.57 ms).
static int pdl_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ECFFT_B200_PDL");
    v = e ? atoi(e) : 1;
    if (v < 0 || v > 2) v = 1;
  }
  return v;
}
static bool pdl_for(size_t tiles) { return pdl_mode() == 2 || (pdl_mode() == 1 && (tiles <= 64 || tiles >= 2048)); }
// ECFFT_B200_TMA (default 1): tile loads as cp.async.bulk.tensor copies completing on an mbarrier; 0 = cp.async.
// Measured on one B200, same box, bit-identical results (profiles/r02_m_ab_tma.txt): ENTER 2^22 14.27 -> 13.95 ms,
// 2^19 2.097 -> 2.070 ms, EXTEND 2^20 0.356 -> 0.350 ms, EXIT 2^22 31.19 -> 31.08 ms (its strided-view passes keep cp.async).
static int tma_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ECFFT_B200_TMA");
    v = e ? atoi(e) : 1;
  }
  return v;
}
// ECFFT_B200_TW_PREFETCH: 0 = off (default), 1 = next stage's twiddles into L1, 2 = into L2.  Measured slower
// (ENTER 2^22 14.87 -> 15.09 / 15.07 ms, profiles/r02_j_ab_tma_prefetch_enter22.txt): the loads it hides were
// already covered by the other resident CTAs, and the address arithmetic costs issue slots.
static uint32_t tw_prefetch_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ECFFT_B200_TW_PREFETCH");
    v = e ? atoi(e) : 0;
    if (v < 0 || v > 2) v = 0;
  }
  return (uint32_t)v;
}
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qr;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qr) != cudaSuccess || qr != cudaDriverEntryPointSuccess) f = nullptr;
    return (EncodeTiledFn)f;
  }();
  return fn;
}
// Tensor map of a pass's input: [total >> rs][2^rs][8 x u32], box {4, 2^min(log_c, 8), 2^krows}; false when the pass
// does not have that shape (batches that are not whole rows, tiny tiles): the caller keeps cp.async.
static bool make_tile_map(const SymParams& p, CUtensorMap* map) {
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc || p.kind != 0 || p.in_shift > 4) return false;
  const uint32_t rs = p.packed ? p.log_t : p.row_shift;
  const uint32_t lc = p.packed ? p.log_t : p.log_c, kr = p.packed ? 0u : p.krows;
  if (p.log_t < 9) return false;   // shared-memory box addresses must be 128-byte aligned; tiny tiles keep cp.async
  if (p.total < ((unsigned long long)1 << p.log_t) || (p.total & (((unsigned long long)1 << rs) - 1))) return false;
  if ((reinterpret_cast<uintptr_t>(p.in) & 15) || rs > 26) return false;
  const uint32_t bc_log = lc < 8 ? lc : 8;
  const cuuint64_t dims[3] = {8, (cuuint64_t)1 << rs, (cuuint64_t)(p.total >> rs)};
  // bytes, dimensions 1 and 2; a strided view (REDC reads every second element, src/fftree.rs:233) is a larger element pitch
  const cuuint64_t strides[2] = {(cuuint64_t)sizeof(Fp) << p.in_shift, (cuuint64_t)sizeof(Fp) << (rs + p.in_shift)};
  const cuuint32_t box[3] = {4, 1u << bc_log, lc > 8 ? 1u : 1u << kr};
  const cuuint32_t estr[3] = {1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, const_cast<Fp*>(p.in + p.in_off), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <class... Args>
static void launch_kernel(void (*kern)(Args...), size_t tiles, int nt, size_t smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)tiles);
  cfg.blockDim = dim3(nt);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_for(tiles) ? 1 : 0;
  ECFFT_CUDA(cudaLaunchKernelEx(&cfg, kern, args...));
}
template <int NT, int MINB>
static void launch_shape(const SymParams& p, size_t tiles, cudaStream_t st) {
  static PerDeviceOnce configured;
  configured.run([] {
    ECFFT_CUDA(cudaFuncSetAttribute(k_extend_sym<NT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((1u << 11) * sizeof(Fp))));
    ECFFT_CUDA(cudaFuncSetAttribute(k_extend_sym_tma<NT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((1u << 11) * sizeof(Fp))));
  });
  const size_t smem = ((size_t)sizeof(Fp)) << p.log_t;
  if (tma_enabled()) {
    alignas(64) CUtensorMap map;
    if (make_tile_map(p, &map)) {
      launch_kernel(k_extend_sym_tma<NT, MINB>, tiles, NT, smem, st, p, map);
      return;
    }
  }
  launch_kernel(k_extend_sym<NT, MINB>, tiles, NT, smem, st, p);
}

// algorithmic bytes (level-streaming model of the reference algorithm): every level reads and writes each
// element once (64 B) and reads its 2^j matrices (128 B) once; a combine moves 128 B per element
static double pass_alg_bytes(const SymParams& p) {
  if (p.kind == 1) return 128.0 * (double)p.total;
  const double levels = (double)(p.lvl_hi - p.lvl_lo) * (p.do_d + p.do_r);
  double mats = 0;
  for (uint32_t j = p.lvl_lo; j < p.lvl_hi; j++) mats += (double)(p.do_d + p.do_r) * 128.0 * (double)(1ull << j);
  return levels * 64.0 * (double)p.total + mats + (p.comb ? 128.0 * (double)p.total : 0.0);
}

static void launch_sym(const SymParams& p_in, cudaStream_t st) {
  SymParams p = p_in;
  p.pf = tw_prefetch_mode();
  p.tma_fence = tma_enabled() == 2;
  // default off: measured 2 % SLOWER at every size (ENTER 2^22 13.96 -> 14.26 ms, EXIT 31.1 -> 31.9 ms; profiles/r02_r_ab_l2pf.txt)
  static const int l2pf = [] { const char* e = getenv("ECFFT_B200_L2PF"); return e ? atoi(e) : 0; }();
  p.l2pf = (uint32_t)l2pf;
  const size_t tiles = (p.total + ((size_t)1 << p.log_t) - 1) >> p.log_t;
  if (tiles > 0x7fffffffull) throw Error(ERR_INVALID_ARG, "extend: grid too large");
  const bool timed = prof::enabled();
  if (timed) prof::record_begin(prof::EXTEND_TILE, pass_alg_bytes(p), st);
  // Shapes 0 and 1 run the same 1024-element tile with 128 threads, so the choice is made per launch (sym_auto_shape):
  // five CTAs of 96 registers per SM for multi-wave grids, four of 122 registers otherwise.
  int shape = sym_variant();
  if (shape == 1 && sym_auto_shape() && tiles >= 2048) shape = 0;
  switch (shape) {
    case 1: launch_shape<128, 4>(p, tiles, st); break;
    case 2: launch_shape<256, 3>(p, tiles, st); break;
    case 3: launch_shape<256, 2>(p, tiles, st); break;
    case 4: launch_shape<64, 10>(p, tiles, st); break;
    default: launch_shape<128, 5>(p, tiles, st); break;
  }
  if (timed) prof::record_end(st);
  prof::count_launch();
  ECFFT_CUDA(cudaGetLastError());
}

// All passes of the symmetric EXTEND of nvec vectors of length 2^log_h, appended to flow.passes.  comb != null
// fuses ENTER's combine into the last pass (nvec even: vectors 2w, 2w+1 are u, v of block w); returns false
// when this depth cannot be fused (the caller then runs EXTEND and the combine separately).
bool plan_extend_sym(SymFlow& flow, const Fp* tw_d, const Fp* tw_r, const Fp* ctr, const Fp* in, Fp* out, uint32_t log_h, size_t nvec,
                     const Fp* pre, const Fp* post, const SymCombine* comb, const SymIO* io) {
  std::vector<SymParams>& L = flow.passes;
  const uint32_t LT = sym_log_tile();
  const size_t total = nvec << log_h;
  if (total < 4 || log_h == 0 || log_h > 31) return false;
  if (comb && (nvec & 1)) return false;
  if (comb && io) throw Error(ERR_INVALID_ARG, "extend_sym: strided views and the fused combine exclude each other");
  SymParams p{};
  // strided views apply to the first load and the last store; passes in between use a contiguous buffer
  SymParams first_io{}, last_io{};
  Fp* mid = out;
  if (io) {
    first_io.in_shift = io->in_shift; first_io.in_off = io->in_off;
    last_io.out_shift = io->out_shift; last_io.out_off = io->out_off;
    last_io.E = io->E; last_io.Z = io->Z; last_io.e_shift = io->e_shift; last_io.e_off = io->e_off;
    last_io.split = io->split;
    if (io->E && !(io->Z && post)) throw Error(ERR_INVALID_ARG, "extend_sym: E needs Z and post");
    mid = io->work;
  }
  auto set_first = [&](SymParams& q) { q.in_shift = first_io.in_shift; q.in_off = first_io.in_off; };
  auto clear_first = [&](SymParams& q) { q.in_shift = 0; q.in_off = 0; };
  auto set_last = [&](SymParams& q) {
    q.out_shift = last_io.out_shift; q.out_off = last_io.out_off;
    q.E = last_io.E; q.Z = last_io.Z; q.e_shift = last_io.e_shift; q.e_off = last_io.e_off;
    q.split = last_io.split;
  };
  p.tw_d = tw_d;
  p.tw_r = tw_r;
  p.ctr = ctr;
  p.total = total;
  p.log_h = log_h;
  p.cond = io ? io->cond : nullptr;
  if (comb) {
    p.A = comb->A;
    p.xnn = comb->xnn;
    p.gam = comb->gam;
    p.gx = comb->gx;
    p.ce0 = comb->e0;
    p.ce1 = comb->e1;
  }
  const uint32_t need = log_h + (comb ? 1 : 0);  // tile bits that hold a whole vector (pair)
  if (need <= LT) {
    // whole vectors fit a tile: 2^(log_t-log_h) consecutive vectors per CTA, the entire EXTEND in one launch
    p.in = in;
    p.out = comb ? comb->out : out;
    p.packed = 1;
    p.log_t = need < 2 ? 2 : need;
    while (p.log_t < LT && ((size_t)1 << p.log_t) < total) p.log_t++;
    p.lvl_lo = 0; p.lvl_hi = log_h; p.boff = 0; p.do_d = 1; p.do_r = 1;
    p.pre = pre;
    p.post = comb ? nullptr : post;
    p.comb = comb ? 1 : 0;
    p.nv = 1;
    set_first(p);
    set_last(p);
    L.push_back(p);
    return true;
  }
  if (io && !mid) throw Error(ERR_INVALID_ARG, "extend_sym: multi-pass EXTEND with strided views needs a work buffer");
  if (log_h < LT) return false;                     // only reachable for comb with need == LT + 1
  const uint32_t outer = log_h - LT;
  if (comb && outer == 0) return false;             // h == tile: no room for the sibling vector
  const uint32_t kmax = LT - 5;                     // strided tiles keep rows of >= 32 (pair: 16) contiguous elements
  const uint32_t npass = (outer + kmax - 1) / kmax;
  std::vector<uint32_t> bounds;                     // level boundaries from log_h down to LT
  bounds.push_back(log_h);
  for (uint32_t i = 1; i <= npass; i++) bounds.push_back(log_h - (outer * i) / npass);
  p.log_t = LT;
  const Fp* src = in;
  for (uint32_t i = 0; i < npass; i++) {            // outer decompose passes, top levels first
    p.in = src; p.out = mid;
    if (i == 0) set_first(p); else clear_first(p);
    p.packed = 0; p.pair = 0; p.comb = 0;
    p.lvl_hi = bounds[i]; p.lvl_lo = bounds[i + 1];
    p.krows = p.lvl_hi - p.lvl_lo; p.log_c = LT - p.krows; p.row_shift = p.lvl_lo; p.boff = p.log_c - p.lvl_lo;
    p.nv = nvec;
    p.do_d = 1; p.do_r = 0;
    p.pre = i == 0 ? pre : nullptr;
    p.post = nullptr;
    L.push_back(p);
    src = mid;
  }
  // inner pass: all levels below the tile size on contiguous tiles
  p.in = src; p.out = npass == 0 ? out : mid;
  if (npass == 0) { set_first(p); set_last(p); } else clear_first(p);
  p.packed = 1; p.pair = 0; p.comb = 0; p.nv = 1;
  p.lvl_lo = 0; p.lvl_hi = LT; p.boff = 0; p.do_d = 1; p.do_r = 1;
  p.pre = npass == 0 ? pre : nullptr;
  p.post = npass == 0 ? post : nullptr;
  L.push_back(p);
  for (uint32_t i = npass; i-- > 0;) {              // outer recombine passes, top levels last
    const bool fin = i == 0;
    p.in = mid; p.out = (fin && comb) ? comb->out : (fin ? out : mid);
    if (fin) set_last(p);
    p.packed = 0;
    p.pair = (fin && comb) ? 1 : 0;
    p.comb = p.pair;
    p.lvl_hi = bounds[i]; p.lvl_lo = bounds[i + 1];
    p.krows = p.lvl_hi - p.lvl_lo; p.log_c = LT - p.pair - p.krows; p.row_shift = p.lvl_lo; p.boff = p.log_c - p.lvl_lo;
    p.nv = p.pair ? nvec / 2 : nvec;
    p.do_d = 0; p.do_r = 1;
    p.pre = nullptr;
    p.post = (fin && !comb) ? post : nullptr;
    L.push_back(p);
  }
  return true;
}

// One launch per pass (the per-pass kernel).  Same result as running the planned passes as a flow.
bool extend_sym(const Fp* tw_d, const Fp* tw_r, const Fp* ctr, const Fp* in, Fp* out, uint32_t log_h, size_t nvec, const Fp* pre, const Fp* post,
                const SymCombine* comb, cudaStream_t st, const SymIO* io) {
  SymFlow flow;
  if (!plan_extend_sym(flow, tw_d, tw_r, ctr, in, out, log_h, nvec, pre, post, comb, io)) return false;
  for (const SymParams& p : flow.passes) launch_sym(p, st);
  return true;
}

// ENTER's combine as a pass of its own (depths whose EXTEND cannot carry it: h = 1, where EXTEND is the
// identity, and h = tile).  W = [u1 | v1] per block, unscaled (gam / gx carry Gamma^1).
void plan_combine_only(SymFlow& flow, const SymCombine& c, const Fp* W, uint32_t log_h, size_t n) {
  SymParams p{};
  p.kind = 1;
  p.in = W;
  p.out = c.out;
  p.A = c.A;
  p.xnn = c.xnn;
  p.gam = c.gam;
  p.gx = c.gx;
  p.total = n;
  p.log_h = log_h;
  p.log_t = sym_log_tile();
  p.nv = 1;
  flow.passes.push_back(p);
}

static int flow_order() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("ECFFT_B200_FLOW_ORDER");
    v = e ? atoi(e) : 0;
  }
  return v;
}
static unsigned long long* g_flow_stats = nullptr;   // device buffer of 4 accumulators when diagnostics are on
void flow_stats_enable(bool on) {
  if (on && !g_flow_stats) {
    ECFFT_CUDA(cudaMalloc((void**)&g_flow_stats, 4 * sizeof(unsigned long long)));
    ECFFT_CUDA(cudaMemset(g_flow_stats, 0, 4 * sizeof(unsigned long long)));
  } else if (!on && g_flow_stats) {
    cudaFree(g_flow_stats);
    g_flow_stats = nullptr;
  }
}
void flow_stats_read(unsigned long long out4[4]) {
  if (!g_flow_stats) throw Error(ERR_INVALID_ARG, "flow statistics are not enabled");
  ECFFT_CUDA(cudaDeviceSynchronize());
  ECFFT_CUDA(cudaMemcpy(out4, g_flow_stats, 4 * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
  ECFFT_CUDA(cudaMemset(g_flow_stats, 0, 4 * sizeof(unsigned long long)));
}
static inline uint32_t umax(uint32_t a, uint32_t b) { return a > b ? a : b; }
static inline uint32_t umin(uint32_t a, uint32_t b) { return a < b ? a : b; }

// Queue order, dependencies and counters of a flow.  Pass k reads, per tile, inside ONE aligned block of
// 2^Din elements of its input and writes inside one aligned block of 2^Dout elements of its output:
//   packed tile      Din = Dout = log_t            (whole vectors or a slice of one)
//   strided tile     Din = Dout = lvl_hi           (rows 2^lvl_lo apart span a 2^lvl_hi block of the vector)
//   pair tile        Din = Dout = log_h + 1        (u and v are neighbours; the outputs interleave over the block)
//   combine-only     Din = Dout = max(log_t, log_h + 1)
// The counters of pass k have the granularity G_k = max(Dout_k, Din_{k+1}); a tile of pass k+1 waits for the
// one counter covering its input block to reach the number of tiles of pass k in such a block.  Inside a
// pass the queue runs block-major over the PREVIOUS pass's counters — (block, vector, tile in block) — so
// tiles are taken in the order their inputs complete.  Returns the number of counters.
static uint32_t finalize_flow(std::vector<SymParams>& P) {
  const size_t K = P.size();
  std::vector<uint32_t> Din(K), Dout(K), G(K);
  for (size_t k = 0; k < K; k++) {
    const SymParams& p = P[k];
    if (p.total % ((size_t)1 << p.log_t)) throw Error(ERR_INVALID_ARG, "flow: a pass does not tile evenly");
    uint32_t dd;
    if (p.kind == 1) dd = umax(p.log_t, p.log_h + 1);
    else if (p.packed) dd = p.log_t;
    else if (p.pair) dd = p.log_h + 1;
    else dd = p.lvl_hi;
    Din[k] = Dout[k] = dd;
  }
  uint32_t ncounters = 0, tiles = 0;
  for (size_t k = 0; k < K; k++) {
    SymParams& p = P[k];
    G[k] = k + 1 < K ? umax(Dout[k], Din[k + 1]) : Dout[k];
    if (p.total % ((size_t)1 << G[k])) throw Error(ERR_INVALID_ARG, "flow: a pass does not split into whole dependency blocks");
    p.sig_base = ncounters;
    p.sig_shift = G[k];
    ncounters += (uint32_t)(p.total >> G[k]);
    p.ntiles = (uint32_t)(p.total >> p.log_t);
    p.tile_begin = tiles;
    tiles += p.ntiles;
    if (k == 0) {
      p.dep_need = 0;
      p.dep_base = p.dep_shift = 0;
    } else {
      p.dep_base = P[k - 1].sig_base;
      p.dep_shift = G[k - 1];
      p.dep_need = 1u << (G[k - 1] - P[k - 1].log_t);
    }
    // order
    const uint32_t gin = k == 0 ? Din[0] : G[k - 1];
    uint32_t tv_log;  // log2 tiles per ordering vector
    if (p.kind == 1 || (p.packed && p.log_h <= p.log_t)) {
      p.ord_nv = 1; p.ord_log_v = 0; tv_log = 0;      // natural order over the whole batch
      p.ord_tpb_log = 0;
    } else {
      const uint32_t log_v = p.pair ? p.log_h + 1 : p.log_h;
      const size_t nv = p.total >> log_v;
      if (nv == 0 || nv > 0xffffffffull) throw Error(ERR_INVALID_ARG, "flow: bad batch");
      p.ord_nv = (uint32_t)nv;
      p.ord_log_v = log_v;
      tv_log = log_v - p.log_t;
      // vector-major (natural) order by default: a pair's two vectors then complete together, half-way
      // through a pass, and a tile's producers sit about one whole pass earlier in the queue.
      // ECFFT_B200_FLOW_ORDER=1: block-major over the previous pass's counters (measured slower).
      p.ord_tpb_log = flow_order() == 1 ? (gin > p.log_t ? umin(gin - p.log_t, tv_log) : 0) : tv_log;
    }
  }
  return ncounters;
}

template <int NT, int MINB>
static void launch_flow_shape(const FlowDesc& d, uint32_t log_t, uint32_t widest_pass, cudaStream_t st) {
  static PerDeviceOnce configured;
  static int slots_per_sm[64];
  int dev = 0;
  ECFFT_CUDA(cudaGetDevice(&dev));
  const size_t smem = ((size_t)sizeof(Fp)) << log_t;
  configured.run([&] {
    ECFFT_CUDA(cudaFuncSetAttribute(k_sym_flow<NT, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((1u << 11) * sizeof(Fp))));
    int nb = 0, sms = 0;
    ECFFT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, k_sym_flow<NT, MINB>, NT, smem));
    ECFFT_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    if (dev < 64) slots_per_sm[dev] = (nb > 0 ? nb : 1) * sms;
  });
  unsigned grid = dev < 64 && slots_per_sm[dev] > 0 ? (unsigned)slots_per_sm[dev] : 148u * MINB;
  // no more CTAs than the widest pass has tiles: the others could only take tiles of later passes and spin
  if (grid > widest_pass) grid = widest_pass;
  if (grid > d.total_tiles) grid = d.total_tiles;
  k_sym_flow<NT, MINB><<<grid, NT, smem, st>>>(d);
}

void launch_flow(SymFlow& flow, cudaStream_t st) {
  if (flow.passes.empty()) return;
  if (flow.passes.size() > (size_t)FLOW_MAX_PASSES) throw Error(ERR_INVALID_ARG, "flow: too many passes");
  const uint32_t ncounters = finalize_flow(flow.passes);
  FlowDesc d{};
  uint32_t log_t = 0, widest = 1;
  double bytes = 0;
  for (size_t k = 0; k < flow.passes.size(); k++) {
    d.pass[k] = flow.passes[k];
    log_t = umax(log_t, flow.passes[k].log_t);
    widest = umax(widest, flow.passes[k].ntiles);
    bytes += pass_alg_bytes(flow.passes[k]);
  }
  flow.alg_bytes = bytes;
  d.npass = (uint32_t)flow.passes.size();
  const SymParams& last = flow.passes.back();
  const unsigned long long tt = (unsigned long long)last.tile_begin + last.ntiles;
  if (tt > 0x7fffffffull) throw Error(ERR_INVALID_ARG, "flow: too many tiles");
  d.total_tiles = (uint32_t)tt;
  unsigned int* mem = nullptr;
  ECFFT_CUDA(cudaMallocAsync((void**)&mem, ((size_t)ncounters + 1) * sizeof(unsigned int), st));
  ECFFT_CUDA(cudaMemsetAsync(mem, 0, ((size_t)ncounters + 1) * sizeof(unsigned int), st));
  d.queue = mem;
  d.counters = mem + 1;
  d.stats = g_flow_stats;
  const bool timed = prof::enabled();
  if (timed) prof::record_begin(prof::EXTEND_TILE, bytes, st);
  switch (sym_variant()) {
    case 1: launch_flow_shape<128, 4>(d, log_t, widest, st); break;
    case 2: launch_flow_shape<256, 3>(d, log_t, widest, st); break;
    case 3: launch_flow_shape<256, 2>(d, log_t, widest, st); break;
    case 4: launch_flow_shape<64, 10>(d, log_t, widest, st); break;
    default: launch_flow_shape<128, 5>(d, log_t, widest, st); break;
  }
  if (timed) prof::record_end(st);
  prof::count_launch();
  cudaError_t e = cudaGetLastError();
  cudaFreeAsync(mem, st);
  ECFFT_CUDA(e);
}

}  // namespace k
}  // namespace ecfft
