// secp256k1 base-field arithmetic for sm_100a (and, for unit tests, the host).
//
// Replaces the arithmetic the reference gets from ark-ff 0.4
// `Fp256<MontBackend<FqConfig,4>>` (reference src/lib.rs:31-37; call sites
// src/utils.rs:338-347 and src/fftree.rs:157-158,187-189,217-219,238,253-255).
//
// Representation on device: 8 x u32 little-endian limbs, value in [0, 2^256)
// ("lazy": congruent mod p, not necessarily < p) inside kernels, canonical
// (< p) whenever a value is stored to global memory.
//
// Two multipliers are provided and benchmarked (tools/microbench.cu):
//   * fp_mul / MulAcc + fp_reduce : schoolbook 8x8 u32 product built from
//     mad.lo.cc / madc.hi.cc carry chains (ptxas fuses each lo/hi pair into
//     one IMAD.WIDE.U32.X), followed by a pseudo-Mersenne fold using
//     p = 2^256 - 0x1000003D1.  Several products can be accumulated into the
//     same 512-bit accumulator before ONE reduction (lazy reduction of the
//     2x2 mat-vec row).  This is x*y mod p with no Montgomery factor.
//   * fp_mont_mul : CIOS Montgomery (a*b*R^-1 mod p, R = 2^256), the same
//     function ark-ff computes, kept as the measured alternative.
//
// Why the plain multiplier can serve a Montgomery-form API: every table
// constant c is kept in plain form on device, data x~ = x*R arrives in
// Montgomery form; c * x~ mod p = (c*x)*R = Montgomery form of c*x, which is
// exactly what mont_mul(c~, x~) yields.  Field elements have one canonical bit
// pattern, so results are bit-identical.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define FP_HD __host__ __device__ __forceinline__
#define FP_D __device__ __forceinline__
#else
#define FP_HD inline
#define FP_D inline
#endif

namespace ecfft {

// ---------------------------------------------------------------------------
// carry-flag primitives: PTX on device, emulated on host (unit tests only)
// ---------------------------------------------------------------------------
#if !defined(__CUDA_ARCH__)
static thread_local uint32_t fp_host_cf = 0;
#endif

FP_HD uint32_t add_cc(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  uint32_t r; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
  uint64_t t = (uint64_t)a + b; fp_host_cf = (uint32_t)(t >> 32); return (uint32_t)t;
#endif
}
FP_HD uint32_t addc_cc(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  uint32_t r; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
  uint64_t t = (uint64_t)a + b + fp_host_cf; fp_host_cf = (uint32_t)(t >> 32); return (uint32_t)t;
#endif
}
FP_HD uint32_t addc(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  uint32_t r; asm volatile("addc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
  return a + b + fp_host_cf;
#endif
}
FP_HD uint32_t sub_cc(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  uint32_t r; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
  uint64_t t = (uint64_t)a - b; fp_host_cf = (uint32_t)(t >> 63); return (uint32_t)t;
#endif
}
FP_HD uint32_t subc_cc(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  uint32_t r; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
  uint64_t t = (uint64_t)a - b - fp_host_cf; fp_host_cf = (uint32_t)(t >> 63); return (uint32_t)t;
#endif
}
FP_HD uint32_t subc(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  uint32_t r; asm volatile("subc.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b)); return r;
#else
  return a - b - fp_host_cf;
#endif
}
FP_HD uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
FP_HD uint32_t mul_hi(uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
  return __umulhi(a, b);
#else
  return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
FP_HD uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
#ifdef __CUDA_ARCH__
  uint32_t r; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
  uint64_t t = (uint64_t)(uint32_t)(a * b) + c; fp_host_cf = (uint32_t)(t >> 32); return (uint32_t)t;
#endif
}
FP_HD uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) {
#ifdef __CUDA_ARCH__
  uint32_t r; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
  uint64_t t = (uint64_t)(uint32_t)(a * b) + c + fp_host_cf; fp_host_cf = (uint32_t)(t >> 32); return (uint32_t)t;
#endif
}
FP_HD uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
#ifdef __CUDA_ARCH__
  uint32_t r; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
  uint64_t t = (((uint64_t)a * b) >> 32) + c; fp_host_cf = (uint32_t)(t >> 32); return (uint32_t)t;
#endif
}
FP_HD uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) {
#ifdef __CUDA_ARCH__
  uint32_t r; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
  uint64_t t = (((uint64_t)a * b) >> 32) + c + fp_host_cf; fp_host_cf = (uint32_t)(t >> 32); return (uint32_t)t;
#endif
}
FP_HD uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) {
#ifdef __CUDA_ARCH__
  uint32_t r; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c)); return r;
#else
  return (uint32_t)(((uint64_t)a * b) >> 32) + c + fp_host_cf;
#endif
}

// ---------------------------------------------------------------------------
// constants: p = 2^256 - DELTA, DELTA = 2^32 + 977
// ---------------------------------------------------------------------------
#define FP_C977 977u
#define FP_P0 0xFFFFFC2Fu
#define FP_P1 0xFFFFFFFEu
#define FP_PX 0xFFFFFFFFu
#define FP_MONT_NP0 0xD2253531u  // -p^-1 mod 2^32

struct alignas(16) Fp {
  uint32_t v[8];
};

FP_HD Fp fp_zero() { Fp r; for (int i = 0; i < 8; i++) r.v[i] = 0; return r; }
FP_HD Fp fp_one() { Fp r = fp_zero(); r.v[0] = 1; return r; }
FP_HD uint32_t fp_p_limb(int i) { return i == 0 ? FP_P0 : (i == 1 ? FP_P1 : FP_PX); }

FP_HD bool fp_eq(const Fp& a, const Fp& b) {
  uint32_t d = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) d |= a.v[i] ^ b.v[i];
  return d == 0;
}
FP_HD bool fp_is_zero(const Fp& a) {
  uint32_t d = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) d |= a.v[i];
  return d == 0;
}

// canonical representative: x in [0,2^256) -> x mod p  (x < 2p always holds)
FP_HD Fp fp_canon(const Fp& x) {
  // x >= p  <=>  x + DELTA overflows 2^256
  Fp s;
  s.v[0] = add_cc(x.v[0], FP_C977);
  s.v[1] = addc_cc(x.v[1], 1u);
#pragma unroll
  for (int i = 2; i < 8; i++) s.v[i] = addc_cc(x.v[i], 0u);
  uint32_t c = addc(0u, 0u);
  Fp r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = c ? s.v[i] : x.v[i];
  return r;
}

// (a + b) mod p for canonical a, b; result canonical
FP_HD Fp fp_add(const Fp& a, const Fp& b) {
  Fp s;
  s.v[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) s.v[i] = addc_cc(a.v[i], b.v[i]);
  uint32_t c = addc(0u, 0u);
  // t = s + DELTA ; if (c || carry(t)) result = t (mod 2^256) else s
  Fp t;
  t.v[0] = add_cc(s.v[0], FP_C977);
  t.v[1] = addc_cc(s.v[1], 1u);
#pragma unroll
  for (int i = 2; i < 8; i++) t.v[i] = addc_cc(s.v[i], 0u);
  uint32_t c2 = addc(0u, 0u);
  uint32_t sel = c | c2;
  Fp r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = sel ? t.v[i] : s.v[i];
  return r;
}

// (a - b) mod p for canonical a, b; result canonical
FP_HD Fp fp_sub(const Fp& a, const Fp& b) {
  Fp d;
  d.v[0] = sub_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) d.v[i] = subc_cc(a.v[i], b.v[i]);
  uint32_t borrow = subc(0u, 0u);  // 0 or 0xFFFFFFFF
  // if borrow: d += p  <=> d -= DELTA (mod 2^256)
  uint32_t m977 = borrow & FP_C977, m1 = borrow & 1u;
  Fp r;
  r.v[0] = sub_cc(d.v[0], m977);
  r.v[1] = subc_cc(d.v[1], m1);
#pragma unroll
  for (int i = 2; i < 8; i++) r.v[i] = subc_cc(d.v[i], 0u);
  return r;
}

FP_HD Fp fp_neg(const Fp& a) { return fp_sub(fp_zero(), a); }

// a + b mod p for lazy a, b in [0,2^256); lazy result.  A carry out of 2^256 folds back as +DELTA
// (the wrapped sum is < 2^256 - 1, so adding DELTA can carry at most once more, which is folded too).
FP_HD Fp fp_add_lazy(const Fp& a, const Fp& b) {
  Fp s;
  s.v[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) s.v[i] = addc_cc(a.v[i], b.v[i]);
  uint32_t c = addc(0u, 0u);
  uint32_t m = 0u - c;
  s.v[0] = add_cc(s.v[0], m & FP_C977);
  s.v[1] = addc_cc(s.v[1], c);
#pragma unroll
  for (int i = 2; i < 8; i++) s.v[i] = addc_cc(s.v[i], 0u);
  uint32_t c2 = addc(0u, 0u);  // only if the wrapped sum was >= 2^256 - DELTA: then the result is tiny
  uint32_t m2 = 0u - c2;
  s.v[0] = add_cc(s.v[0], m2 & FP_C977);
  s.v[1] = addc_cc(s.v[1], c2);
  s.v[2] = addc(s.v[2], 0u);
  return s;
}

// a - b mod p for lazy a in [0,2^256) and CANONICAL b; lazy result in [0,2^256)
FP_HD Fp fp_sub_lazy(const Fp& a, const Fp& b) {
  Fp d;
  d.v[0] = sub_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) d.v[i] = subc_cc(a.v[i], b.v[i]);
  uint32_t borrow = subc(0u, 0u);
  // a < b <= p-1: a - b + p lies in (0, p): one correction is exact
  uint32_t m977 = borrow & FP_C977, m1 = borrow & 1u;
  Fp r;
  r.v[0] = sub_cc(d.v[0], m977);
  r.v[1] = subc_cc(d.v[1], m1);
#pragma unroll
  for (int i = 2; i < 8; i++) r.v[i] = subc_cc(d.v[i], 0u);
  return r;
}

// a - b mod p for lazy a AND lazy b in [0,2^256); lazy result.  a - b = d - borrow*2^256 with
// d the wrapped difference, and 2^256 = DELTA (mod p): subtract DELTA once per borrow.  The first
// correction can borrow again only when d < DELTA; the wrapped value then lies in
// [2^256 - DELTA, 2^256), whose limbs >= 2 are all ones and whose low 64 bits exceed 2*DELTA,
// so the second correction touches two limbs and cannot borrow.
FP_HD Fp fp_sub_lazy2(const Fp& a, const Fp& b) {
  Fp d;
  d.v[0] = sub_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) d.v[i] = subc_cc(a.v[i], b.v[i]);
  uint32_t b1 = subc(0u, 0u);  // 0 or 0xFFFFFFFF
  Fp r;
  r.v[0] = sub_cc(d.v[0], b1 & FP_C977);
  r.v[1] = subc_cc(d.v[1], b1 & 1u);
#pragma unroll
  for (int i = 2; i < 8; i++) r.v[i] = subc_cc(d.v[i], 0u);
  uint32_t b2 = subc(0u, 0u);
  r.v[0] = sub_cc(r.v[0], b2 & FP_C977);
  r.v[1] = subc(r.v[1], b2 & 1u);
  return r;
}

// Short forms of fp_add_lazy / fp_sub_lazy2 for the butterfly kernels: the DELTA correction is applied
// to the low two limbs only and the (probability ~2^-31) carry/borrow out of them takes a branch.
FP_HD Fp fp_add_lazy_f(const Fp& a, const Fp& b) {
#ifdef FP_BRANCHFREE_ADDSUB   // A/B switch: the straight-line forms (more instructions, no basic-block boundary inside a butterfly)
  return fp_add_lazy(a, b);
#endif
  Fp s;
  s.v[0] = add_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) s.v[i] = addc_cc(a.v[i], b.v[i]);
  uint32_t c = addc(0u, 0u);
  s.v[0] = add_cc(s.v[0], (0u - c) & FP_C977);
  s.v[1] = addc_cc(s.v[1], c);
  uint32_t c2 = addc(0u, 0u);
  if (c2) {  // ripple into the high limbs; a second wrap leaves a tiny value that takes one more DELTA
    s.v[2] = add_cc(s.v[2], 1u);
#pragma unroll
    for (int i = 3; i < 8; i++) s.v[i] = addc_cc(s.v[i], 0u);
    uint32_t c3 = addc(0u, 0u);
    s.v[0] = add_cc(s.v[0], (0u - c3) & FP_C977);
    s.v[1] = addc_cc(s.v[1], c3);
    s.v[2] = addc(s.v[2], 0u);
  }
  return s;
}
FP_HD Fp fp_sub_lazy2_f(const Fp& a, const Fp& b) {
#ifdef FP_BRANCHFREE_ADDSUB
  return fp_sub_lazy2(a, b);
#endif
  Fp d;
  d.v[0] = sub_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < 8; i++) d.v[i] = subc_cc(a.v[i], b.v[i]);
  uint32_t b1 = subc(0u, 0u);  // 0 or 0xFFFFFFFF
  d.v[0] = sub_cc(d.v[0], b1 & FP_C977);
  d.v[1] = subc_cc(d.v[1], b1 & 1u);
  uint32_t b2 = subc(0u, 0u);
  if (b2) {  // borrow ripples through the high limbs; if it falls off the top, one more DELTA (cannot borrow)
    d.v[2] = sub_cc(d.v[2], 1u);
#pragma unroll
    for (int i = 3; i < 8; i++) d.v[i] = subc_cc(d.v[i], 0u);
    uint32_t b3 = subc(0u, 0u);
    d.v[0] = sub_cc(d.v[0], b3 & FP_C977);
    d.v[1] = subc(d.v[1], b3 & 1u);
  }
  return d;
}

// ---------------------------------------------------------------------------
// 512(+1)-bit product accumulator, split into an even-aligned and an
// odd-aligned half so every partial product lands on a 64-bit-aligned limb
// pair (=> IMAD.WIDE) with the carry riding the CC chain.
// value = sum e[k] 2^(32k) + sum o[k] 2^(32(k+1))
// ---------------------------------------------------------------------------
struct MulAcc {
  uint32_t e[17];
  uint32_t o[15];
};

FP_HD void acc_clear(MulAcc& A) {
#pragma unroll
  for (int i = 0; i < 17; i++) A.e[i] = 0;
#pragma unroll
  for (int i = 0; i < 15; i++) A.o[i] = 0;
}
// start the accumulator at a 256-bit addend (for x0 + k*x1 style fused ops)
FP_HD void acc_set(MulAcc& A, const Fp& x) {
  acc_clear(A);
#pragma unroll
  for (int i = 0; i < 8; i++) A.e[i] = x.v[i];
}

// w[0..7] += {a[0],a[2],a[4],a[6]} * b ; carry-out added into w[8]
FP_HD void mad_row4(uint32_t* w, const uint32_t* a, uint32_t b) {
  w[0] = mad_lo_cc(a[0], b, w[0]);
  w[1] = madc_hi_cc(a[0], b, w[1]);
  w[2] = madc_lo_cc(a[2], b, w[2]);
  w[3] = madc_hi_cc(a[2], b, w[3]);
  w[4] = madc_lo_cc(a[4], b, w[4]);
  w[5] = madc_hi_cc(a[4], b, w[5]);
  w[6] = madc_lo_cc(a[6], b, w[6]);
  w[7] = madc_hi_cc(a[6], b, w[7]);
  w[8] = addc(w[8], 0u);
}

// one row i of the product a*b (multiplier limb b.v[i]) into the accumulator
template <int I>
FP_HD void acc_row(MulAcc& A, const Fp& a, const Fp& b) {
  if ((I & 1) == 0) {
    mad_row4(&A.e[I], &a.v[0], b.v[I]);
    mad_row4(&A.o[I], &a.v[1], b.v[I]);
  } else {
    mad_row4(&A.o[I - 1], &a.v[0], b.v[I]);
    mad_row4(&A.e[I + 1], &a.v[1], b.v[I]);
  }
}

// A += a*b
FP_HD void acc_mul(MulAcc& A, const Fp& a, const Fp& b) {
  acc_row<0>(A, a, b); acc_row<1>(A, a, b); acc_row<2>(A, a, b); acc_row<3>(A, a, b);
  acc_row<4>(A, a, b); acc_row<5>(A, a, b); acc_row<6>(A, a, b); acc_row<7>(A, a, b);
}

// A += a*b + c*d, rows interleaved so that a row's carry-out never ripples
// further than one limb (see DESIGN.md "carry bound")
FP_HD void acc_mul2(MulAcc& A, const Fp& a, const Fp& b, const Fp& c, const Fp& d) {
  acc_row<0>(A, a, b); acc_row<0>(A, c, d);
  acc_row<1>(A, a, b); acc_row<1>(A, c, d);
  acc_row<2>(A, a, b); acc_row<2>(A, c, d);
  acc_row<3>(A, a, b); acc_row<3>(A, c, d);
  acc_row<4>(A, a, b); acc_row<4>(A, c, d);
  acc_row<5>(A, a, b); acc_row<5>(A, c, d);
  acc_row<6>(A, a, b); acc_row<6>(A, c, d);
  acc_row<7>(A, a, b); acc_row<7>(A, c, d);
}

// merge the two halves and fold 2^256 = DELTA (mod p) until < 2^256.
// Input value must be < 2^514 (true for <= 3 accumulated products + addend).
FP_HD Fp fp_reduce(const MulAcc& A) {
  uint32_t t[17];
  t[0] = A.e[0];
  t[1] = add_cc(A.e[1], A.o[0]);
#pragma unroll
  for (int k = 2; k < 16; k++) t[k] = addc_cc(A.e[k], A.o[k - 1]);
  t[16] = addc(A.e[16], 0u);

#ifdef FP_REDUCE_CHAINED   // the round-1 default: three carry chains on one array — 5 % faster in the radix-2 tile kernel
                          // (profiles/r01_g_ab_variants.txt), 2 % SLOWER in k_extend_sym than the form below, whose chains
                          // never change register-pair alignment: ptxas places the moves the chained form needs on the
                          // multiplier's own pipe as IMAD.MOV (655 -> 369 moves per kernel; ENTER 2^22 14.13 -> 13.84 ms,
                          // EXIT 2^22 30.59 -> 30.04 ms, profiles/r02_aj_ab_arith_variants.txt)
  // r = lo + 977*hi + (hi << 32), hi = t[8..16]  (r < 2^291 -> 10 limbs): three chains on one array
  uint32_t r[10];
  r[0] = mad_lo_cc(t[8], FP_C977, t[0]);
  r[1] = madc_hi_cc(t[8], FP_C977, t[1]);
  r[2] = madc_lo_cc(t[10], FP_C977, t[2]);
  r[3] = madc_hi_cc(t[10], FP_C977, t[3]);
  r[4] = madc_lo_cc(t[12], FP_C977, t[4]);
  r[5] = madc_hi_cc(t[12], FP_C977, t[5]);
  r[6] = madc_lo_cc(t[14], FP_C977, t[6]);
  r[7] = madc_hi_cc(t[14], FP_C977, t[7]);
  r[8] = madc_lo_cc(t[16], FP_C977, 0u);
  r[9] = addc(0u, 0u);
  r[1] = mad_lo_cc(t[9], FP_C977, r[1]);
  r[2] = madc_hi_cc(t[9], FP_C977, r[2]);
  r[3] = madc_lo_cc(t[11], FP_C977, r[3]);
  r[4] = madc_hi_cc(t[11], FP_C977, r[4]);
  r[5] = madc_lo_cc(t[13], FP_C977, r[5]);
  r[6] = madc_hi_cc(t[13], FP_C977, r[6]);
  r[7] = madc_lo_cc(t[15], FP_C977, r[7]);
  r[8] = madc_hi_cc(t[15], FP_C977, r[8]);
  r[9] = addc(r[9], 0u);
  r[1] = add_cc(r[1], t[8]);
#pragma unroll
  for (int k = 2; k < 9; k++) r[k] = addc_cc(r[k], t[k + 7]);
  r[9] = addc(r[9], t[16]);

#else
  // r = lo + 977*hi + (hi << 32), hi = t[8..16]  (r < 2^291 -> 10 limbs).
  // Even-aligned chain: lo + 977*{t8,t10,t12,t14,t16}; odd-aligned chain (weight 2^32):
  // 977*{t9,t11,t13,t15} + (hi << 32) taken as the 64-bit addends (t8,t9),(t10,t11),... — both chains
  // write fresh aligned register pairs, so no limb has to change pair alignment (no moves).
  uint32_t ev[10], od[9];
  ev[0] = mad_lo_cc(t[8], FP_C977, t[0]);
  ev[1] = madc_hi_cc(t[8], FP_C977, t[1]);
  ev[2] = madc_lo_cc(t[10], FP_C977, t[2]);
  ev[3] = madc_hi_cc(t[10], FP_C977, t[3]);
  ev[4] = madc_lo_cc(t[12], FP_C977, t[4]);
  ev[5] = madc_hi_cc(t[12], FP_C977, t[5]);
  ev[6] = madc_lo_cc(t[14], FP_C977, t[6]);
  ev[7] = madc_hi_cc(t[14], FP_C977, t[7]);
  ev[8] = madc_lo_cc(t[16], FP_C977, 0u);
  ev[9] = addc(0u, 0u);
  od[0] = mad_lo_cc(t[9], FP_C977, t[8]);
  od[1] = madc_hi_cc(t[9], FP_C977, t[9]);
  od[2] = madc_lo_cc(t[11], FP_C977, t[10]);
  od[3] = madc_hi_cc(t[11], FP_C977, t[11]);
  od[4] = madc_lo_cc(t[13], FP_C977, t[12]);
  od[5] = madc_hi_cc(t[13], FP_C977, t[13]);
  od[6] = madc_lo_cc(t[15], FP_C977, t[14]);
  od[7] = madc_hi_cc(t[15], FP_C977, t[15]);
  od[8] = addc(t[16], 0u);
  uint32_t r[10];
  r[0] = ev[0];
  r[1] = add_cc(ev[1], od[0]);
#pragma unroll
  for (int k = 2; k < 9; k++) r[k] = addc_cc(ev[k], od[k - 1]);
  r[9] = addc(ev[9], od[8]);

#endif
  // second fold: h2 = r[8] + r[9]*2^32 (< 2^35); V = h2*DELTA < 2^68
  uint32_t v0 = mul_lo(r[8], FP_C977);
  uint32_t v1 = mul_hi(r[8], FP_C977) + r[9] * FP_C977;  // < 2^10 + 2^13
  v1 = add_cc(v1, r[8]);
  uint32_t v2 = addc(r[9], 0u);
  Fp x;
  x.v[0] = add_cc(r[0], v0);
  x.v[1] = addc_cc(r[1], v1);
  x.v[2] = addc_cc(r[2], v2);
#pragma unroll
  for (int k = 3; k < 8; k++) x.v[k] = addc_cc(r[k], 0u);
  uint32_t c = addc(0u, 0u);
  // third fold: if the add wrapped, x (now < 2^68) += DELTA; cannot wrap again
  uint32_t m = 0u - c;
  x.v[0] = add_cc(x.v[0], m & FP_C977);
  x.v[1] = addc_cc(x.v[1], c);
  x.v[2] = addc(x.v[2], 0u);
  return x;
}

// a*b mod p, lazy result in [0,2^256)
FP_HD Fp fp_mul_lazy(const Fp& a, const Fp& b) {
  MulAcc A; acc_clear(A); acc_mul(A, a, b); return fp_reduce(A);
}
FP_HD Fp fp_mul(const Fp& a, const Fp& b) { return fp_canon(fp_mul_lazy(a, b)); }
FP_HD Fp fp_sqr(const Fp& a) { return fp_mul(a, a); }
// a*b + c*d mod p with one reduction (lazy)
FP_HD Fp fp_dot2_lazy(const Fp& a, const Fp& b, const Fp& c, const Fp& d) {
  MulAcc A; acc_clear(A); acc_mul2(A, a, b, c, d); return fp_reduce(A);
}
// x + a*b mod p with one reduction (lazy)
FP_HD Fp fp_muladd_lazy(const Fp& x, const Fp& a, const Fp& b) {
  MulAcc A; acc_set(A, x); acc_mul(A, a, b); return fp_reduce(A);
}

// ---------------------------------------------------------------------------
// CIOS Montgomery multiplication, R = 2^256 (what ark-ff computes); the
// measured alternative to fp_mul.  Inputs/outputs canonical Montgomery form.
// ---------------------------------------------------------------------------
FP_HD Fp fp_mont_mul(const Fp& a, const Fp& b) {
  uint32_t t[10];
#pragma unroll
  for (int i = 0; i < 10; i++) t[i] = 0;
#pragma unroll
  for (int i = 0; i < 8; i++) {
    // t += a * b[i]
    uint32_t carry = 0;
#pragma unroll
    for (int j = 0; j < 8; j++) {
      uint32_t lo = mad_lo_cc(a.v[j], b.v[i], t[j]);
      uint32_t hi = madc_hi(a.v[j], b.v[i], 0u);
      t[j] = add_cc(lo, carry);
      carry = addc(hi, 0u);
    }
    t[8] = add_cc(t[8], carry);
    t[9] = addc(0u, 0u);
    // m = t[0] * np0 ; t = (t + m*p) >> 32
    uint32_t m = t[0] * FP_MONT_NP0;
    uint32_t lo = mad_lo_cc(m, FP_P0, t[0]);
    carry = madc_hi(m, FP_P0, 0u);
    (void)lo;
#pragma unroll
    for (int j = 1; j < 8; j++) {
      uint32_t pj = (j == 1) ? FP_P1 : FP_PX;
      uint32_t l2 = mad_lo_cc(m, pj, t[j]);
      uint32_t h2 = madc_hi(m, pj, 0u);
      t[j - 1] = add_cc(l2, carry);
      carry = addc(h2, 0u);
    }
    t[7] = add_cc(t[8], carry);
    t[8] = addc(t[9], 0u);
  }
  // conditional subtract p
  Fp s;
  s.v[0] = sub_cc(t[0], FP_P0);
  s.v[1] = subc_cc(t[1], FP_P1);
#pragma unroll
  for (int i = 2; i < 8; i++) s.v[i] = subc_cc(t[i], FP_PX);
  uint32_t borrow = subc(t[8], 0u);  // top limb after borrow: 0xFFFFFFFF if t < p
  Fp r;
#pragma unroll
  for (int i = 0; i < 8; i++) r.v[i] = (borrow >> 31) ? t[i] : s.v[i];
  return r;
}

// ---------------------------------------------------------------------------
// Montgomery <-> plain conversion constants (plain-form values)
//   R      = 2^256 mod p = DELTA
//   R^-1 mod p
// to_mont(x)   = x * R      (plain multiply by DELTA)
// from_mont(x) = x * R^-1   (plain multiply by RINV)
// ---------------------------------------------------------------------------
FP_HD Fp fp_const_R() { Fp r = fp_zero(); r.v[0] = FP_C977; r.v[1] = 1u; return r; }
FP_HD Fp fp_const_RINV() {
  // (2^256)^-1 mod p, verified in tests/test_fp_host.py against Python pow()
  Fp r;
  r.v[0] = 0x0868192Au; r.v[1] = 0xD838091Du; r.v[2] = 0xDC24A059u; r.v[3] = 0xBCB223FEu;
  r.v[4] = 0x95F2B761u; r.v[5] = 0x9C46C2C2u; r.v[6] = 0x15538399u; r.v[7] = 0xC9BD1905u;
  return r;
}

// x^e for a 64-bit exponent (square-and-multiply, MSB first); canonical
FP_HD Fp fp_pow_u64(const Fp& x, uint64_t e) {
  Fp r = fp_one();
  bool started = false;
  for (int i = 63; i >= 0; i--) {
    if (started) r = fp_mul_lazy(r, r);
    if ((e >> i) & 1) {
      r = started ? fp_mul_lazy(r, x) : x;
      started = true;
    }
  }
  return fp_canon(r);
}

// n squarings
FP_HD Fp fp_sqr_n(Fp x, int n) {
  for (int i = 0; i < n; i++) x = fp_mul_lazy(x, x);
  return x;
}
// x^(p-2) (Fermat inverse); 0 -> 0.  p-2 in binary is 223 ones, a zero, 22 ones, 0000, 1, 0, 11, 0, 1:
// blocks of ones of length {1,2,22,223} -> addition chain with 255 squarings + 15 multiplications.
FP_HD Fp fp_inv(const Fp& a) {
  Fp x2 = fp_mul_lazy(fp_sqr_n(a, 1), a);
  Fp x3 = fp_mul_lazy(fp_sqr_n(x2, 1), a);
  Fp x6 = fp_mul_lazy(fp_sqr_n(x3, 3), x3);
  Fp x9 = fp_mul_lazy(fp_sqr_n(x6, 3), x3);
  Fp x11 = fp_mul_lazy(fp_sqr_n(x9, 2), x2);
  Fp x22 = fp_mul_lazy(fp_sqr_n(x11, 11), x11);
  Fp x44 = fp_mul_lazy(fp_sqr_n(x22, 22), x22);
  Fp x88 = fp_mul_lazy(fp_sqr_n(x44, 44), x44);
  Fp x176 = fp_mul_lazy(fp_sqr_n(x88, 88), x88);
  Fp x220 = fp_mul_lazy(fp_sqr_n(x176, 44), x44);
  Fp x223 = fp_mul_lazy(fp_sqr_n(x220, 3), x3);
  Fp t = fp_mul_lazy(fp_sqr_n(x223, 23), x22);
  t = fp_mul_lazy(fp_sqr_n(t, 5), a);
  t = fp_mul_lazy(fp_sqr_n(t, 3), x2);
  t = fp_mul_lazy(fp_sqr_n(t, 2), a);
  return fp_canon(t);
}
// x^((p+1)/4): square root candidate for p = 3 (mod 4) (ark-ff Case3Mod4); caller checks r*r == x.
// (p+1)/4 = 2^254 - 2^30 - 244: 223 ones, 0, 22 ones, 0000, 11, 00
FP_HD Fp fp_sqrt_candidate(const Fp& a) {
  Fp x2 = fp_mul_lazy(fp_sqr_n(a, 1), a);
  Fp x3 = fp_mul_lazy(fp_sqr_n(x2, 1), a);
  Fp x6 = fp_mul_lazy(fp_sqr_n(x3, 3), x3);
  Fp x9 = fp_mul_lazy(fp_sqr_n(x6, 3), x3);
  Fp x11 = fp_mul_lazy(fp_sqr_n(x9, 2), x2);
  Fp x22 = fp_mul_lazy(fp_sqr_n(x11, 11), x11);
  Fp x44 = fp_mul_lazy(fp_sqr_n(x22, 22), x22);
  Fp x88 = fp_mul_lazy(fp_sqr_n(x44, 44), x44);
  Fp x176 = fp_mul_lazy(fp_sqr_n(x88, 88), x88);
  Fp x220 = fp_mul_lazy(fp_sqr_n(x176, 44), x44);
  Fp x223 = fp_mul_lazy(fp_sqr_n(x220, 3), x3);
  Fp t = fp_mul_lazy(fp_sqr_n(x223, 23), x22);
  t = fp_mul_lazy(fp_sqr_n(t, 6), x2);
  t = fp_sqr_n(t, 2);
  return fp_canon(t);
}

#if defined(__CUDACC__)
// 32-byte element moved as two 16-byte vector accesses
FP_D Fp fp_load(const Fp* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = q[0], b = q[1];
  Fp r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
FP_D Fp fp_load_ro(const Fp* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldg(q), b = __ldg(q + 1);
  Fp r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
// L2-only load: data another CTA of the same launch may have written (never served from L1)
FP_D Fp fp_load_cg(const Fp* p) {
  const uint4* q = reinterpret_cast<const uint4*>(p);
  uint4 a = __ldcg(q), b = __ldcg(q + 1);
  Fp r;
  r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
  r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
  return r;
}
FP_D void fp_store(Fp* p, const Fp& x) {
  uint4* q = reinterpret_cast<uint4*>(p);
  q[0] = make_uint4(x.v[0], x.v[1], x.v[2], x.v[3]);
  q[1] = make_uint4(x.v[4], x.v[5], x.v[6], x.v[7]);
}
#endif

}  // namespace ecfft
