// Affine arithmetic on the ECFFT-II "Good Curve" E_{a,B}: y^2 = x^3 + a x^2 + B x, B = b^2
// (reference src/ec.rs:27-35, 142-173, 376-424), plain-form field values, host + device.
#pragma once
#include "fp.cuh"

namespace ecfft {

struct Pt {
  Fp x, y;
  bool inf;
};
FP_HD Pt pt_infinity() {
  Pt r;
  r.x = fp_zero();
  r.y = fp_zero();
  r.inf = true;
  return r;
}
// Point + Point, reference src/ec.rs:376-424 with a1 = a3 = a6 = 0, a2 = a, a4 = B.
// lambda and nu share a denominator, so one inversion serves both.
FP_HD Pt pt_add(const Pt& p, const Pt& q, const Fp& a, const Fp& a4) {
  if (p.inf) return q;
  if (q.inf) return p;
  bool same_x = fp_eq(p.x, q.x);
  if (same_x && fp_is_zero(fp_add(p.y, q.y))) return pt_infinity();
  Fp lnum, nnum, den;
  if (same_x) {  // tangent
    Fp xx = fp_sqr(p.x), ax = fp_mul(a, p.x);
    lnum = fp_add(fp_add(fp_add(xx, xx), xx), fp_add(fp_add(ax, ax), a4));
    nnum = fp_add(fp_neg(fp_mul(xx, p.x)), fp_mul(a4, p.x));
    den = fp_add(p.y, p.y);
  } else {  // chord
    lnum = fp_sub(q.y, p.y);
    nnum = fp_sub(fp_mul(p.y, q.x), fp_mul(q.y, p.x));
    den = fp_sub(q.x, p.x);
  }
  Fp dinv = fp_inv(den);
  Fp lambda = fp_mul(lnum, dinv), nu = fp_mul(nnum, dinv);
  Pt r;
  r.inf = false;
  r.x = fp_sub(fp_sub(fp_sub(fp_sqr(lambda), a), p.x), q.x);
  r.y = fp_sub(fp_neg(fp_mul(lambda, r.x)), nu);
  return r;
}

}  // namespace ecfft
