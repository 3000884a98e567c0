/* ORACLE — TEST INFRASTRUCTURE ONLY.  See ecfft_oracle.h for scope, citations and pinning.
 * CPU restatement of andrewmilson/ecfft (reference @ 9ca932a) for secp256k1::Fp. */
#include "ecfft_oracle.h"

#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef unsigned __int128 u128;

/* ------------------------------------------------------------------------- */
/* Field: ark-ff 0.4 Fp256<MontBackend<FqConfig,4>> restated (reference       */
/* src/lib.rs:31-37 declares it; the arithmetic lives in ark-ff).             */
/* ------------------------------------------------------------------------- */
static const uint64_t P[4] = {0xFFFFFFFEFFFFFC2FULL, 0xFFFFFFFFFFFFFFFFULL, 0xFFFFFFFFFFFFFFFFULL,
                              0xFFFFFFFFFFFFFFFFULL};
static const uint64_t NP0 = 0xd838091dd2253531ULL;            /* -p^-1 mod 2^64 */
static const fe FE_ONE = {{0x00000001000003d1ULL, 0, 0, 0}};  /* R mod p */
static const fe FE_R2 = {{0x000007a2000e90a1ULL, 1, 0, 0}};   /* R^2 mod p */
static const fe FE_ZERO = {{0, 0, 0, 0}};

static int fe_is_zero(const fe* a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static int fe_eq(const fe* a, const fe* b) {
  return a->l[0] == b->l[0] && a->l[1] == b->l[1] && a->l[2] == b->l[2] && a->l[3] == b->l[3];
}
static int geq_p(const uint64_t t[4]) {
  for (int i = 3; i >= 0; i--) {
    if (t[i] > P[i]) return 1;
    if (t[i] < P[i]) return 0;
  }
  return 1;
}
static void sub_p(uint64_t t[4]) {
  u128 b = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)t[i] - P[i] - b;
    t[i] = (uint64_t)d;
    b = (d >> 64) & 1;
  }
}

/* Montgomery multiplication, coarsely integrated operand scanning, 4 x 64-bit limbs */
static void fe_mul(fe* r, const fe* a, const fe* b) {
  uint64_t t[6] = {0, 0, 0, 0, 0, 0};
  for (int i = 0; i < 4; i++) {
    u128 c = 0;
    for (int j = 0; j < 4; j++) {
      c += (u128)a->l[j] * b->l[i] + t[j];
      t[j] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[4] = (uint64_t)c;
    t[5] = (uint64_t)(c >> 64);
    uint64_t m = t[0] * NP0;
    c = (u128)m * P[0] + t[0];
    c >>= 64;
    for (int j = 1; j < 4; j++) {
      c += (u128)m * P[j] + t[j];
      t[j - 1] = (uint64_t)c;
      c >>= 64;
    }
    c += t[4];
    t[3] = (uint64_t)c;
    t[4] = t[5] + (uint64_t)(c >> 64);
  }
  if (t[4] || geq_p(t)) sub_p(t);
  memcpy(r->l, t, 32);
}
static void fe_sqr(fe* r, const fe* a) { fe_mul(r, a, a); }
static void fe_add(fe* r, const fe* a, const fe* b) {
  uint64_t t[4];
  u128 c = 0;
  for (int i = 0; i < 4; i++) {
    c += (u128)a->l[i] + b->l[i];
    t[i] = (uint64_t)c;
    c >>= 64;
  }
  if (c || geq_p(t)) sub_p(t);
  memcpy(r->l, t, 32);
}
static void fe_sub(fe* r, const fe* a, const fe* b) {
  uint64_t t[4];
  u128 bo = 0;
  for (int i = 0; i < 4; i++) {
    u128 d = (u128)a->l[i] - b->l[i] - bo;
    t[i] = (uint64_t)d;
    bo = (d >> 64) & 1;
  }
  if (bo) {
    u128 c = 0;
    for (int i = 0; i < 4; i++) {
      c += (u128)t[i] + P[i];
      t[i] = (uint64_t)c;
      c >>= 64;
    }
  }
  memcpy(r->l, t, 32);
}
static void fe_neg(fe* r, const fe* a) { fe_sub(r, &FE_ZERO, a); }
static void fe_dbl(fe* r, const fe* a) { fe_add(r, a, a); }

/* x^e, e given as 4 little-endian u64 limbs (square-and-multiply; value is path independent) */
static void fe_pow4(fe* r, const fe* x, const uint64_t e[4]) {
  fe acc = FE_ONE;
  int started = 0;
  for (int l = 3; l >= 0; l--)
    for (int i = 63; i >= 0; i--) {
      if (started) fe_sqr(&acc, &acc);
      if ((e[l] >> i) & 1) {
        fe_mul(&acc, &acc, x);
        started = 1;
      }
    }
  *r = acc;
}
static void fe_pow(fe* r, const fe* x, uint64_t e) {
  uint64_t ee[4] = {e, 0, 0, 0};
  fe_pow4(r, x, ee);
}
/* Field::inverse — ark-ff uses a binary Euclid variant; the inverse is unique, so x^(p-2) */
static int fe_inv(fe* r, const fe* x) {
  if (fe_is_zero(x)) return 0;
  uint64_t e[4] = {P[0] - 2, P[1], P[2], P[3]};
  fe_pow4(r, x, e);
  return 1;
}
/* Field::sqrt for p = 3 (mod 4): ark-ff SqrtPrecomputation::Case3Mod4, x^((p+1)/4), checked */
static int fe_sqrt(fe* r, const fe* x) {
  uint64_t q[4] = {P[0] + 1, P[1], P[2], P[3]}; /* p+1 (no carry: P[0] ends in ...2F) */
  uint64_t e[4];
  for (int i = 0; i < 4; i++) e[i] = (q[i] >> 2) | (i < 3 ? q[i + 1] << 62 : 0);
  fe s, s2;
  fe_pow4(&s, x, e);
  fe_sqr(&s2, &s);
  if (!fe_eq(&s2, x)) return 0;
  *r = s;
  return 1;
}
/* ark_ff::batch_inversion: Montgomery's trick, zero entries are left untouched */
static void batch_inversion(fe* v, size_t n) {
  if (n == 0) return;
  fe* prod = (fe*)malloc(n * sizeof(fe));
  fe acc = FE_ONE;
  for (size_t i = 0; i < n; i++) {
    if (!fe_is_zero(&v[i])) fe_mul(&acc, &acc, &v[i]);
    prod[i] = acc;
  }
  fe inv;
  fe_inv(&inv, &acc); /* acc != 0 */
  for (size_t i = n; i-- > 0;) {
    if (fe_is_zero(&v[i])) continue;
    fe prev = FE_ONE;
    for (size_t j = i; j-- > 0;)
      if (!fe_is_zero(&v[j])) {
        prev = prod[j];
        break;
      }
    fe new_inv;
    fe_mul(&new_inv, &inv, &v[i]);
    fe_mul(&v[i], &inv, &prev);
    inv = new_inv;
  }
  free(prod);
}
/* same as above but O(n) when there are no zeros in between (the common case) */
static void batch_inversion_fast(fe* v, size_t n) {
  for (size_t i = 0; i < n; i++)
    if (fe_is_zero(&v[i])) {
      batch_inversion(v, n);
      return;
    }
  if (n == 0) return;
  fe* prod = (fe*)malloc(n * sizeof(fe));
  fe acc = FE_ONE;
  for (size_t i = 0; i < n; i++) {
    prod[i] = acc; /* product of v[0..i) */
    fe_mul(&acc, &acc, &v[i]);
  }
  fe inv;
  fe_inv(&inv, &acc);
  for (size_t i = n; i-- > 0;) {
    fe new_inv;
    fe_mul(&new_inv, &inv, &v[i]);
    fe_mul(&v[i], &inv, &prod[i]);
    inv = new_inv;
  }
  free(prod);
}

/* Thread budget of the tree BUILD (orc_set_build_threads; default 1 = the reference's single thread).  The
 * loops it cuts into ranges are element-wise, so the tables are bit-identical for any budget; it only
 * shortens the set-up of the CPU baseline at n = 2^22 (bench.py --impl reference). */
static int g_build_threads = 1;
void orc_set_build_threads(int threads) { g_build_threads = threads < 1 ? 1 : threads; }
static void parallel_for(size_t n, int threads, void (*fn)(size_t, size_t, void*), void* ctx);

void orc_fe_mul(const fe* a, const fe* b, fe* r) { fe_mul(r, a, b); }
void orc_fe_add(const fe* a, const fe* b, fe* r) { fe_add(r, a, b); }
void orc_fe_sub(const fe* a, const fe* b, fe* r) { fe_sub(r, a, b); }
void orc_fe_inv(const fe* a, fe* r) {
  if (!fe_inv(r, a)) *r = FE_ZERO;
}
void orc_batch_inversion(fe* v, size_t n) { batch_inversion_fast(v, n); }
void orc_fe_from_canonical(const uint8_t b[32], fe* r) {
  fe x;
  for (int i = 0; i < 4; i++) {
    uint64_t w = 0;
    for (int k = 7; k >= 0; k--) w = (w << 8) | b[8 * i + k];
    x.l[i] = w;
  }
  fe_mul(r, &x, &FE_R2);
}
void orc_fe_to_canonical(const fe* a, uint8_t b[32]) {
  fe one = {{1, 0, 0, 0}}, x;
  fe_mul(&x, a, &one);
  for (int i = 0; i < 4; i++)
    for (int k = 0; k < 8; k++) b[8 * i + k] = (uint8_t)(x.l[i] >> (8 * k));
}
static fe fe_from_hex(const char* hex) { /* big-endian hex string, canonical -> Montgomery */
  uint8_t b[32];
  memset(b, 0, 32);
  size_t n = strlen(hex);
  for (size_t i = 0; i < n; i++) {
    char ch = hex[n - 1 - i];
    uint8_t d = (ch >= '0' && ch <= '9') ? ch - '0' : (ch >= 'a' && ch <= 'f') ? ch - 'a' + 10 : ch - 'A' + 10;
    b[i / 2] |= (i & 1) ? d << 4 : d;
  }
  fe r;
  orc_fe_from_canonical(b, &r);
  return r;
}

/* ------------------------------------------------------------------------- */
/* RationalMap (reference src/utils.rs:367-390) with DensePolynomial coeffs    */
/* ------------------------------------------------------------------------- */
typedef struct {
  size_t nnum, nden;
  fe* num;
  fe* den;
} ratmap;

static void poly_eval(fe* r, const fe* c, size_t n, const fe* x) { /* Horner, DensePolynomial::evaluate */
  fe acc = FE_ZERO;
  for (size_t i = n; i-- > 0;) {
    fe_mul(&acc, &acc, x);
    fe_add(&acc, &acc, &c[i]);
  }
  *r = acc;
}
static ratmap ratmap_new(const fe* num, size_t nnum, const fe* den, size_t nden) {
  /* from_coefficients_slice drops trailing zero coefficients */
  while (nnum && fe_is_zero(&num[nnum - 1])) nnum--;
  while (nden && fe_is_zero(&den[nden - 1])) nden--;
  ratmap m;
  m.nnum = nnum;
  m.nden = nden;
  m.num = (fe*)malloc((nnum ? nnum : 1) * sizeof(fe));
  m.den = (fe*)malloc((nden ? nden : 1) * sizeof(fe));
  memcpy(m.num, num, nnum * sizeof(fe));
  memcpy(m.den, den, nden * sizeof(fe));
  return m;
}
static ratmap ratmap_clone(const ratmap* s) { return ratmap_new(s->num, s->nnum, s->den, s->nden); }
static void ratmap_free(ratmap* m) {
  free(m->num);
  free(m->den);
}
/* RationalMap::map, utils.rs:383-385; returns 0 for None */
static int ratmap_map(const ratmap* m, const fe* x, fe* r) {
  fe nu, de, di;
  poly_eval(&nu, m->num, m->nnum, x);
  poly_eval(&de, m->den, m->nden, x);
  if (!fe_inv(&di, &de)) return 0;
  fe_mul(r, &nu, &di);
  return 1;
}

/* ------------------------------------------------------------------------- */
/* Good curve E_{a,B}: y^2 = x^3 + a x^2 + B x, B = b^2 (reference src/ec.rs)  */
/* ------------------------------------------------------------------------- */
typedef struct { fe a, b; } curve;            /* GoodCurve::Odd, ec.rs:34 */
typedef struct { fe x, y; int inf; curve c; } point; /* Point, ec.rs:363-367; inf = curve None */

static int curve_new_odd(curve* c, const fe* a, const fe* bb) { /* ec.rs:38-45 */
  fe t, u;
  fe_sqr(&t, a);
  fe_dbl(&u, bb);
  fe_dbl(&u, &u);
  fe_sub(&t, &t, &u);
  if (fe_is_zero(bb) || fe_is_zero(&t)) return 0;
  fe b;
  if (!fe_sqrt(&b, bb)) return 0;
  fe_add(&t, a, &b);
  fe_add(&t, &t, &b);
  if (!fe_sqrt(&u, &t)) return 0;
  c->a = *a;
  c->b = b;
  return 1;
}
static point point_zero(void) {
  point p;
  memset(&p, 0, sizeof p);
  p.inf = 1;
  return p;
}
/* Point + Point, ec.rs:376-424 with a1 = a3 = a6 = 0, a2 = a, a4 = b^2 (ec.rs:142-173).
 * lambda and nu share a denominator; one inversion serves both (same field values). */
static point point_add(const point* p, const point* q) {
  if (p->inf) return *q;
  if (q->inf) return *p;
  fe a2 = p->c.a, a4, t, u;
  fe_sqr(&a4, &p->c.b);
  const fe *x1 = &p->x, *y1 = &p->y, *x2 = &q->x, *y2 = &q->y;
  fe_add(&t, y1, y2);
  if (fe_eq(x1, x2) && fe_is_zero(&t)) return point_zero();
  fe lnum, nnum, den, dinv, lambda, nu;
  if (fe_eq(x1, x2)) {
    fe x1x1, a2x1;
    fe_sqr(&x1x1, x1);
    fe_mul(&a2x1, &a2, x1);
    fe_add(&lnum, &x1x1, &x1x1);
    fe_add(&lnum, &lnum, &x1x1);
    fe_add(&lnum, &lnum, &a2x1);
    fe_add(&lnum, &lnum, &a2x1);
    fe_add(&lnum, &lnum, &a4);
    fe_mul(&t, &x1x1, x1);
    fe_neg(&t, &t);
    fe_mul(&u, &a4, x1);
    fe_add(&nnum, &t, &u);
    fe_add(&den, y1, y1);
  } else {
    fe_sub(&lnum, y2, y1);
    fe_mul(&t, y1, x2);
    fe_mul(&u, y2, x1);
    fe_sub(&nnum, &t, &u);
    fe_sub(&den, x2, x1);
  }
  fe_inv(&dinv, &den);
  fe_mul(&lambda, &lnum, &dinv);
  fe_mul(&nu, &nnum, &dinv);
  point r;
  r.inf = 0;
  r.c = p->c;
  fe_sqr(&t, &lambda);
  fe_sub(&t, &t, &a2);
  fe_sub(&t, &t, x1);
  fe_sub(&r.x, &t, x2);
  fe_mul(&t, &lambda, &r.x);
  fe_neg(&t, &t);
  fe_sub(&r.y, &t, &nu);
  return r;
}
/* utils.rs:356-365 */
static int two_adicity(point p) {
  for (int i = 0; i < 2048; i++) {
    if (p.inf) return i;
    p = point_add(&p, &p);
  }
  return -1;
}
typedef struct { curve dom, cod; ratmap r, h; } isogeny; /* g is the zero map for odd fields */
/* GoodCurve::good_isogeny, Odd branch, ec.rs:75-88 */
static int good_isogeny(const curve* c, isogeny* iso) {
  fe bb, t, u, a_prime, b_prime;
  fe_sqr(&bb, &c->b);
  fe_dbl(&t, &c->b);   /* 2b */
  fe_dbl(&u, &t);      /* 4b */
  fe_add(&a_prime, &c->a, &u);
  fe_add(&a_prime, &a_prime, &t);
  fe_mul(&t, &c->a, &c->b);
  fe_dbl(&t, &t);
  fe_dbl(&t, &t);      /* 4ab */
  fe_dbl(&u, &bb);
  fe_dbl(&u, &u);
  fe_dbl(&u, &u);      /* 8b^2 */
  fe_add(&b_prime, &t, &u);
  iso->dom = *c;
  if (!curve_new_odd(&iso->cod, &a_prime, &b_prime)) return 0;
  fe m2b, mbb;
  fe_dbl(&m2b, &c->b);
  fe_neg(&m2b, &m2b);
  fe_neg(&mbb, &bb);
  fe rn[3] = {bb, m2b, FE_ONE}, rd[2] = {FE_ZERO, FE_ONE};
  fe hn[3] = {mbb, FE_ZERO, FE_ONE}, hd[3] = {FE_ZERO, FE_ZERO, FE_ONE};
  iso->r = ratmap_new(rn, 3, rd, 2);
  iso->h = ratmap_new(hn, 3, hd, 3);
  return 1;
}
/* Isogeny::map, ec.rs:344-358 (g = 0) */
static point isogeny_map(const isogeny* iso, const point* p) {
  if (p->inf) return point_zero();
  fe rx, hx;
  if (!ratmap_map(&iso->r, &p->x, &rx) || !ratmap_map(&iso->h, &p->x, &hx)) return point_zero();
  point q;
  q.inf = 0;
  q.c = iso->cod;
  q.x = rx;
  fe_mul(&q.y, &hx, &p->y);
  return q;
}

/* ------------------------------------------------------------------------- */
/* FFTree (reference src/fftree.rs:23-38) and BinaryTree addressing            */
/* (src/utils.rs:228-293): layer i of a tree stored in v[0..len) is            */
/* v[(len/2)>>i .. 2*((len/2)>>i)).                                            */
/* ------------------------------------------------------------------------- */
struct orc_tree {
  size_t n;   /* leaves */
  fe* f;      /* 2n */
  fe* rmat;   /* n matrices, 4 fe each, row major */
  fe* dmat;
  size_t nmaps;
  ratmap* maps;
  fe *xnn_s, *xnn_s_inv;                    /* n */
  fe *z0_s1, *z1_s0, *z0_inv_s1, *z1_inv_s0; /* nz = n/2 */
  fe *z0z0, *z1z1;                          /* nzz = n (0 when n == 1) */
  size_t nz, nzz;
  struct orc_tree* sub;
};

static unsigned ilog2(size_t n) {
  unsigned l = 0;
  while (n >>= 1) l++;
  return l;
}
static int is_pow2(size_t n) { return n && !(n & (n - 1)); }
static fe* fe_alloc(size_t n) { return (fe*)malloc((n ? n : 1) * sizeof(fe)); }

size_t orc_tree_leaves(const orc_tree* t) { return t->n; }
const orc_tree* orc_subtree_with_size(const orc_tree* t, size_t n) {
  if (!is_pow2(n)) return NULL;
  while (t && t->n > n) t = t->sub;
  return (t && t->n == n) ? t : NULL;
}
void orc_tree_free(orc_tree* t) {
  if (!t) return;
  orc_tree_free(t->sub);
  free(t->f); free(t->rmat); free(t->dmat);
  for (size_t i = 0; i < t->nmaps; i++) ratmap_free(&t->maps[i]);
  free(t->maps);
  free(t->xnn_s); free(t->xnn_s_inv); free(t->z0_s1); free(t->z1_s0);
  free(t->z0_inv_s1); free(t->z1_inv_s0); free(t->z0z0); free(t->z1z1);
  free(t);
}
size_t orc_tree_table(const orc_tree* t, const char* name, const fe** ptr) {
#define TBL(s, p, c) if (!strcmp(name, s)) { *ptr = (p); return (c); }
  TBL("f", t->f, 2 * t->n)
  TBL("recombine", t->rmat, 4 * t->n)
  TBL("decompose", t->dmat, 4 * t->n)
  TBL("xnn_s", t->xnn_s, t->n)
  TBL("xnn_s_inv", t->xnn_s_inv, t->n)
  TBL("z0_s1", t->z0_s1, t->nz)
  TBL("z1_s0", t->z1_s0, t->nz)
  TBL("z0_inv_s1", t->z0_inv_s1, t->nz)
  TBL("z1_inv_s0", t->z1_inv_s0, t->nz)
  TBL("z0z0_rem_xnn_s", t->z0z0, t->nzz)
  TBL("z1z1_rem_xnn_s", t->z1z1, t->nzz)
#undef TBL
  *ptr = NULL;
  return 0;
}

/* Mat2x2 * [F;2], utils.rs:338-347 */
static void matvec(const fe* m, const fe* x0, const fe* x1, fe* y0, fe* y1) {
  fe a, b, c, d;
  fe_mul(&a, &m[0], x0);
  fe_mul(&b, &m[1], x1);
  fe_mul(&c, &m[2], x0);
  fe_mul(&d, &m[3], x1);
  fe_add(y0, &a, &b);
  fe_add(y1, &c, &d);
}

typedef struct { const orc_tree* t; const fe* in; size_t n; int moiety; fe* out; int depth; } ext_job;
static void extend_impl(const orc_tree* t, const fe* evals, size_t n, int moiety, fe* out, int depth);
static void* ext_thread(void* p) {
  ext_job* j = (ext_job*)p;
  extend_impl(j->t, j->in, j->n, j->moiety, j->out, j->depth);
  return NULL;
}
/* FFTree::extend_impl, fftree.rs:72-120.  `moiety` is the TARGET.  depth > 0 forks the
 * two independent recursive calls onto threads (orc_enter_mt only). */
static void extend_impl(const orc_tree* t, const fe* evals, size_t n, int moiety, fe* out, int depth) {
  if (n == 1) {
    out[0] = evals[0];
    return;
  }
  unsigned layer = (ilog2(2 * t->n) - 2) - ilog2(n); /* f.num_layers() - 2 - log2 n */
  size_t layer_size = (t->n / 2) >> layer;           /* matrices BinaryTree has n entries */
  size_t h = n / 2;
  fe* e0 = fe_alloc(h);
  fe* e1 = fe_alloc(h);
  const fe* dl = t->dmat + 4 * layer_size;
  for (size_t i = 0; i < h; i++) /* skip(S0 => 1, S1 => 0).step_by(2) */
    matvec(dl + 4 * (2 * i + (moiety == 0 ? 1 : 0)), &evals[i], &evals[i + h], &e0[i], &e1[i]);
  fe* e0p = fe_alloc(h);
  fe* e1p = fe_alloc(h);
  if (depth > 0 && h >= 64) {
    pthread_t th;
    ext_job j = {t, e0, h, moiety, e0p, depth - 1};
    pthread_create(&th, NULL, ext_thread, &j);
    extend_impl(t, e1, h, moiety, e1p, depth - 1);
    pthread_join(th, NULL);
  } else {
    extend_impl(t, e0, h, moiety, e0p, 0);
    extend_impl(t, e1, h, moiety, e1p, 0);
  }
  const fe* rl = t->rmat + 4 * layer_size;
  for (size_t i = 0; i < h; i++) /* skip(S0 => 0, S1 => 1).step_by(2) */
    matvec(rl + 4 * (2 * i + (moiety == 1 ? 1 : 0)), &e0p[i], &e1p[i], &out[i], &out[i + h]);
  free(e0); free(e1); free(e0p); free(e1p);
}
/* fftree.rs:128-135 */
static void mextend_impl(const orc_tree* t, const fe* evals, size_t n, int moiety, fe* out) {
  extend_impl(t, evals, n, moiety, out, 0);
  const fe* z = moiety == 1 ? t->z0_s1 : t->z1_s0;
  for (size_t i = 0; i < n; i++) fe_add(&out[i], &out[i], &z[i]);
}

typedef struct { const orc_tree* t; const fe* in; size_t n; fe* out; int depth; } ent_job;
static void enter_impl(const orc_tree* t, const fe* coeffs, size_t n, fe* out, int depth);
static void* ent_thread(void* p) {
  ent_job* j = (ent_job*)p;
  enter_impl(j->t, j->in, j->n, j->out, j->depth);
  return NULL;
}
/* FFTree::enter_impl, fftree.rs:143-161 */
static void enter_impl(const orc_tree* t, const fe* coeffs, size_t n, fe* out, int depth) {
  if (n == 1) {
    out[0] = coeffs[0];
    return;
  }
  size_t h = n / 2;
  const orc_tree* sub = t->sub;
  fe *u0 = fe_alloc(h), *v0 = fe_alloc(h), *u1 = fe_alloc(h), *v1 = fe_alloc(h);
  if (depth > 0 && h >= 64) {
    pthread_t th;
    ent_job j = {sub, coeffs, h, u0, depth - 1};
    pthread_create(&th, NULL, ent_thread, &j);
    enter_impl(sub, coeffs + h, h, v0, depth - 1);
    pthread_join(th, NULL);
    ext_job e = {t, u0, h, 1, u1, depth - 1};
    pthread_create(&th, NULL, ext_thread, &e);
    extend_impl(t, v0, h, 1, v1, depth - 1);
    pthread_join(th, NULL);
  } else {
    enter_impl(sub, coeffs, h, u0, 0);
    enter_impl(sub, coeffs + h, h, v0, 0);
    extend_impl(t, u0, h, 1, u1, 0);
    extend_impl(t, v0, h, 1, v1, 0);
  }
  for (size_t i = 0; i < h; i++) {
    fe m;
    fe_mul(&m, &v0[i], &t->xnn_s[2 * i]);
    fe_add(&out[2 * i], &u0[i], &m);
    fe_mul(&m, &v1[i], &t->xnn_s[2 * i + 1]);
    fe_add(&out[2 * i + 1], &u1[i], &m);
  }
  free(u0); free(v0); free(u1); free(v1);
}

/* fftree.rs:232-259.  moiety 0 => REDC by Z_0, 1 => by Z_1 */
static void redc_impl(const orc_tree* t, const fe* evals, const fe* a, size_t n, int moiety, fe* out) {
  size_t h = n / 2;
  fe *e0 = fe_alloc(h), *e1 = fe_alloc(h), *a0inv = fe_alloc(h), *a1 = fe_alloc(h);
  for (size_t i = 0; i < h; i++) {
    e0[i] = evals[2 * i];
    e1[i] = evals[2 * i + 1];
    a0inv[i] = a[2 * i];
    a1[i] = a[2 * i + 1];
  }
  batch_inversion_fast(a0inv, h);
  fe* t0 = fe_alloc(h);
  for (size_t i = 0; i < h; i++) fe_mul(&t0[i], &e0[i], &a0inv[i]);
  fe* g1 = fe_alloc(h);
  extend_impl(t, t0, h, moiety == 1 ? 0 : 1, g1, 0);
  const fe* zinv = moiety == 0 ? t->z0_inv_s1 : t->z1_inv_s0;
  fe* h1 = fe_alloc(h);
  for (size_t i = 0; i < h; i++) {
    fe m;
    fe_mul(&m, &g1[i], &a1[i]);
    fe_sub(&m, &e1[i], &m);
    fe_mul(&h1[i], &m, &zinv[i]);
  }
  fe* h0 = fe_alloc(h);
  extend_impl(t, h1, h, moiety, h0, 0);
  for (size_t i = 0; i < h; i++) {
    out[2 * i] = h0[i];
    out[2 * i + 1] = h1[i];
  }
  free(e0); free(e1); free(a0inv); free(a1); free(t0); free(g1); free(h1); free(h0);
}
/* fftree.rs:277-281 */
static void modular_reduce_impl(const orc_tree* t, const fe* evals, const fe* a, const fe* c, size_t n, fe* out) {
  fe* h = fe_alloc(n);
  redc_impl(t, evals, a, n, 0, h);
  for (size_t i = 0; i < n; i++) fe_mul(&h[i], &h[i], &c[i]);
  redc_impl(t, h, a, n, 0, out);
  free(h);
}
/* fftree.rs:200-224 */
static void exit_impl(const orc_tree* t, const fe* evals, size_t n, fe* out) {
  if (n == 1) {
    out[0] = evals[0];
    return;
  }
  size_t h = n / 2;
  fe* m = fe_alloc(n);
  modular_reduce_impl(t, evals, t->xnn_s, t->z0z0, n, m);
  fe *u0 = fe_alloc(h), *v0 = fe_alloc(h);
  for (size_t i = 0; i < h; i++) u0[i] = m[2 * i];
  exit_impl(t->sub, u0, h, out);
  for (size_t i = 0; i < h; i++) {
    fe d;
    fe_sub(&d, &evals[2 * i], &u0[i]);
    fe_mul(&v0[i], &d, &t->xnn_s_inv[2 * i]);
  }
  exit_impl(t->sub, v0, h, out + h);
  free(m); free(u0); free(v0);
}
/* fftree.rs:169-192 */
static size_t degree_impl(const orc_tree* t, const fe* evals, size_t n) {
  if (n == 1) return 0;
  size_t h = n / 2;
  fe *e0 = fe_alloc(h), *e1 = fe_alloc(h), *g1 = fe_alloc(h);
  for (size_t i = 0; i < h; i++) {
    e0[i] = evals[2 * i];
    e1[i] = evals[2 * i + 1];
  }
  extend_impl(t, e0, h, 1, g1, 0);
  int same = 1;
  for (size_t i = 0; i < h && same; i++) same = fe_eq(&g1[i], &e1[i]);
  size_t res;
  if (same) {
    res = degree_impl(t->sub, e0, h);
  } else {
    fe *t1 = fe_alloc(h), *t0 = fe_alloc(h);
    for (size_t i = 0; i < h; i++) {
      fe d;
      fe_sub(&d, &e1[i], &g1[i]);
      fe_mul(&t1[i], &d, &t->z0_inv_s1[i]);
    }
    extend_impl(t, t1, h, 0, t0, 0);
    res = h + degree_impl(t->sub, t0, h);
    free(t1); free(t0);
  }
  free(e0); free(e1); free(g1);
  return res;
}
/* fftree.rs:291-308; output has 2n entries */
static void vanish_impl(const orc_tree* t, const fe* dom, size_t n, fe* out) {
  if (n == 1) {
    const fe* l = t->f + t->n; /* leaves; t->n == 2 here */
    fe_sub(&out[0], &dom[0], &l[0]);
    fe_sub(&out[1], &dom[0], &l[1]);
    return;
  }
  size_t h = n / 2;
  fe *qp = fe_alloc(n), *qpp = fe_alloc(n), *q0 = fe_alloc(n), *q1 = fe_alloc(n);
  vanish_impl(t->sub, dom, h, qp);
  vanish_impl(t->sub, dom + h, h, qpp);
  for (size_t i = 0; i < n; i++) fe_mul(&q0[i], &qp[i], &qpp[i]);
  mextend_impl(t, q0, n, 1, q1); /* self.mextend(&q_s0, S1): subtree_with_size(2n) is self */
  for (size_t i = 0; i < n; i++) {
    out[2 * i] = q0[i];
    out[2 * i + 1] = q1[i];
  }
  free(qp); free(qpp); free(q0); free(q1);
}

/* public wrappers: pick subtree_with_size then call *_impl (fftree.rs:123,138,164,195,227,264,272,286,313) */
int orc_extend(const orc_tree* t, const fe* evals, size_t n, int moiety, fe* out) {
  const orc_tree* s = orc_subtree_with_size(t, n * 2);
  if (!s) return 1;
  extend_impl(s, evals, n, moiety, out, 0);
  return 0;
}
int orc_mextend(const orc_tree* t, const fe* evals, size_t n, int moiety, fe* out) {
  const orc_tree* s = orc_subtree_with_size(t, n * 2);
  if (!s || !s->nz) return 1;
  mextend_impl(s, evals, n, moiety, out);
  return 0;
}
int orc_enter(const orc_tree* t, const fe* coeffs, size_t n, fe* out) {
  const orc_tree* s = orc_subtree_with_size(t, n);
  if (!s) return 1;
  enter_impl(s, coeffs, n, out, 0);
  return 0;
}
/* All-cores variant of enter_impl / extend_impl for the CPU baseline: identical arithmetic and
 * recursion; a thread budget is split between the two independent recursive calls and the butterfly /
 * combine loops are cut into ranges across the budget (plain pthreads; the image has no OpenMP
 * runtime).  The reference library itself is single-threaded. */
typedef struct { void (*fn)(size_t, size_t, void*); size_t lo, hi; void* ctx; } pf_job;
static void* pf_thread(void* p) {
  pf_job* j = (pf_job*)p;
  j->fn(j->lo, j->hi, j->ctx);
  return NULL;
}
static void parallel_for(size_t n, int threads, void (*fn)(size_t, size_t, void*), void* ctx) {
  if (threads > 64) threads = 64;
  if (threads <= 1 || n < 2048) {
    fn(0, n, ctx);
    return;
  }
  pthread_t th[64];
  pf_job jobs[64];
  for (int k = 0; k < threads; k++) {
    jobs[k].fn = fn;
    jobs[k].lo = n * (size_t)k / threads;
    jobs[k].hi = n * (size_t)(k + 1) / threads;
    jobs[k].ctx = ctx;
    if (k < threads - 1) pthread_create(&th[k], NULL, pf_thread, &jobs[k]);
  }
  fn(jobs[threads - 1].lo, jobs[threads - 1].hi, ctx);
  for (int k = 0; k < threads - 1; k++) pthread_join(th[k], NULL);
}
typedef struct { const fe* mats; int skip; const fe *a, *b; fe *y0, *y1; } bf_ctx;
static void bf_range(size_t lo, size_t hi, void* p) {
  bf_ctx* c = (bf_ctx*)p;
  for (size_t i = lo; i < hi; i++) matvec(c->mats + 4 * (2 * i + c->skip), &c->a[i], &c->b[i], &c->y0[i], &c->y1[i]);
}
typedef struct { const orc_tree* t; const fe* in; size_t n; int moiety; fe* out; int threads; } mt_job;
static void extend_mt(const orc_tree* t, const fe* evals, size_t n, int moiety, fe* out, int threads);
static void* extend_mt_thread(void* p) {
  mt_job* j = (mt_job*)p;
  extend_mt(j->t, j->in, j->n, j->moiety, j->out, j->threads);
  return NULL;
}
static void extend_mt(const orc_tree* t, const fe* evals, size_t n, int moiety, fe* out, int threads) {
  if (threads <= 1 || n <= 2048) {
    extend_impl(t, evals, n, moiety, out, 0);
    return;
  }
  unsigned layer = (ilog2(2 * t->n) - 2) - ilog2(n);
  size_t layer_size = (t->n / 2) >> layer;
  size_t h = n / 2;
  fe *e0 = fe_alloc(h), *e1 = fe_alloc(h), *e0p = fe_alloc(h), *e1p = fe_alloc(h);
  bf_ctx d = {t->dmat + 4 * layer_size, moiety == 0 ? 1 : 0, evals, evals + h, e0, e1};
  parallel_for(h, threads, bf_range, &d);
  pthread_t th;
  mt_job j = {t, e0, h, moiety, e0p, threads / 2};
  pthread_create(&th, NULL, extend_mt_thread, &j);
  extend_mt(t, e1, h, moiety, e1p, threads - threads / 2);
  pthread_join(th, NULL);
  bf_ctx r = {t->rmat + 4 * layer_size, moiety == 1 ? 1 : 0, e0p, e1p, out, out + h};
  parallel_for(h, threads, bf_range, &r);
  free(e0); free(e1); free(e0p); free(e1p);
}
typedef struct { const fe *u0, *v0, *u1, *v1, *xnn; fe* out; } cb_ctx;
static void cb_range(size_t lo, size_t hi, void* p) {
  cb_ctx* c = (cb_ctx*)p;
  for (size_t i = lo; i < hi; i++) {
    fe m;
    fe_mul(&m, &c->v0[i], &c->xnn[2 * i]);
    fe_add(&c->out[2 * i], &c->u0[i], &m);
    fe_mul(&m, &c->v1[i], &c->xnn[2 * i + 1]);
    fe_add(&c->out[2 * i + 1], &c->u1[i], &m);
  }
}
static void enter_mt(const orc_tree* t, const fe* coeffs, size_t n, fe* out, int threads);
static void* enter_mt_thread(void* p) {
  mt_job* j = (mt_job*)p;
  enter_mt(j->t, j->in, j->n, j->out, j->threads);
  return NULL;
}
static void enter_mt(const orc_tree* t, const fe* coeffs, size_t n, fe* out, int threads) {
  if (threads <= 1 || n <= 2048) {
    enter_impl(t, coeffs, n, out, 0);
    return;
  }
  size_t h = n / 2;
  fe *u0 = fe_alloc(h), *v0 = fe_alloc(h), *u1 = fe_alloc(h), *v1 = fe_alloc(h);
  pthread_t th;
  mt_job j = {t->sub, coeffs, h, 0, u0, threads / 2};
  pthread_create(&th, NULL, enter_mt_thread, &j);
  enter_mt(t->sub, coeffs + h, h, v0, threads - threads / 2);
  pthread_join(th, NULL);
  mt_job e = {t, u0, h, 1, u1, threads / 2};
  pthread_create(&th, NULL, extend_mt_thread, &e);
  extend_mt(t, v0, h, 1, v1, threads - threads / 2);
  pthread_join(th, NULL);
  cb_ctx c = {u0, v0, u1, v1, t->xnn_s, out};
  parallel_for(h, threads, cb_range, &c);
  free(u0); free(v0); free(u1); free(v1);
}
int orc_enter_mt(const orc_tree* t, const fe* coeffs, size_t n, fe* out, int threads) {
  const orc_tree* s = orc_subtree_with_size(t, n);
  if (!s) return 1;
  enter_mt(s, coeffs, n, out, threads < 1 ? 1 : threads);
  return 0;
}

/* Bottom-up restatement of enter_impl for the recursion depths m_lo < m <= m_hi: `in` holds n/m_lo
 * evaluation vectors of length m_lo (m_lo = 1: coefficients); after the pass for m it holds n/m vectors
 * of length m.  orc_enter_range(t, c, n, 1, n, out) == orc_enter(t, c, n, out); used to check the flat
 * schedule the CUDA path and the multi-GPU split rely on. */
int orc_enter_range(const orc_tree* t, const fe* in, size_t n, size_t m_lo, size_t m_hi, fe* out) {
  if (!is_pow2(n) || !is_pow2(m_lo) || !is_pow2(m_hi) || m_lo > m_hi || m_hi > n) return 1;
  if (!orc_subtree_with_size(t, m_hi)) return 1;
  fe* cur = fe_alloc(n);
  fe* nxt = fe_alloc(n);
  memcpy(cur, in, n * sizeof(fe));
  for (size_t m = 2 * m_lo; m <= m_hi; m *= 2) {
    const orc_tree* s = orc_subtree_with_size(t, m);
    size_t h = m / 2;
    fe *u1 = fe_alloc(h), *v1 = fe_alloc(h);
    for (size_t off = 0; off < n; off += m) {
      const fe *u0 = cur + off, *v0 = cur + off + h;
      extend_impl(s, u0, h, 1, u1, 0);
      extend_impl(s, v0, h, 1, v1, 0);
      for (size_t i = 0; i < h; i++) {
        fe p;
        fe_mul(&p, &v0[i], &s->xnn_s[2 * i]);
        fe_add(&nxt[off + 2 * i], &u0[i], &p);
        fe_mul(&p, &v1[i], &s->xnn_s[2 * i + 1]);
        fe_add(&nxt[off + 2 * i + 1], &u1[i], &p);
      }
    }
    free(u1); free(v1);
    fe* sw = cur; cur = nxt; nxt = sw;
  }
  memcpy(out, cur, n * sizeof(fe));
  free(cur); free(nxt);
  return 0;
}
int orc_exit(const orc_tree* t, const fe* evals, size_t n, fe* out) {
  const orc_tree* s = orc_subtree_with_size(t, n);
  if (!s || (n > 1 && !s->nzz)) return 1;
  exit_impl(s, evals, n, out);
  return 0;
}
int orc_degree(const orc_tree* t, const fe* evals, size_t n, size_t* degree) {
  const orc_tree* s = orc_subtree_with_size(t, n);
  if (!s || (n > 1 && !s->nz)) return 1;
  *degree = degree_impl(s, evals, n);
  return 0;
}
int orc_redc_z0(const orc_tree* t, const fe* evals, const fe* a, size_t n, fe* out) {
  const orc_tree* s = orc_subtree_with_size(t, n);
  if (!s || !s->nz) return 1;
  redc_impl(s, evals, a, n, 0, out);
  return 0;
}
int orc_redc_z1(const orc_tree* t, const fe* evals, const fe* a, size_t n, fe* out) {
  const orc_tree* s = orc_subtree_with_size(t, n);
  if (!s || !s->nz) return 1;
  redc_impl(s, evals, a, n, 1, out);
  return 0;
}
int orc_modular_reduce(const orc_tree* t, const fe* evals, const fe* a, const fe* c, size_t n, fe* out) {
  const orc_tree* s = orc_subtree_with_size(t, n);
  if (!s || !s->nz) return 1;
  modular_reduce_impl(s, evals, a, c, n, out);
  return 0;
}
int orc_vanish(const orc_tree* t, const fe* domain, size_t n, fe* out) {
  const orc_tree* s = orc_subtree_with_size(t, n * 2);
  if (!s || !s->nz) return 1;
  vanish_impl(s, domain, n, out);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* Construction: FFTree::new / from_tree / derive_subtree                      */
/* (fftree.rs:42-70, 318-482)                                                  */
/* ------------------------------------------------------------------------- */
static orc_tree* from_tree(fe* f, size_t n, const ratmap* maps, size_t nmaps, int parts);

/* element-wise loops of from_tree / fftree_new as ranges (see g_build_threads) */
typedef struct { const fe* s; fe *xnnnn, *xnn; uint64_t nnnn, nn; } pow_ctx;
static void pow_range(size_t lo, size_t hi, void* p) {  /* fftree.rs:328-331 */
  pow_ctx* c = (pow_ctx*)p;
  for (size_t i = lo; i < hi; i++) {
    fe_pow(&c->xnnnn[i], &c->s[i], c->nnnn);
    fe_pow(&c->xnn[i], &c->s[i], c->nn);
  }
}
typedef struct { const fe* l; size_t d; const ratmap* map; fe *rl, *dl, *det; } mat_ctx;
static void mat_r_range(size_t lo, size_t hi, void* p) {  /* fftree.rs:354-360 */
  mat_ctx* c = (mat_ctx*)p;
  const size_t d = c->d;
  for (size_t i = lo; i < hi; i++) {
    fe s0 = c->l[i], s1 = c->l[i + d], v0, v1;
    poly_eval(&v0, c->map->den, c->map->nden, &s0);
    poly_eval(&v1, c->map->den, c->map->nden, &s1);
    fe_pow(&v0, &v0, d / 2 - 1);
    fe_pow(&v1, &v1, d / 2 - 1);
    fe* r = c->rl + 4 * i;
    r[0] = v0;
    fe_mul(&r[1], &s0, &v0);
    r[2] = v1;
    fe_mul(&r[3], &s1, &v1);
    /* Mat2x2::inverse, utils.rs:325-335: one determinant inverse per matrix (batched by the caller) */
    fe a, b;
    fe_mul(&a, &r[0], &r[3]);
    fe_mul(&b, &r[1], &r[2]);
    fe_sub(&c->det[i], &a, &b);
  }
}
static void mat_d_range(size_t lo, size_t hi, void* p) {  /* fftree.rs:361 */
  mat_ctx* c = (mat_ctx*)p;
  for (size_t i = lo; i < hi; i++) {
    const fe* r = c->rl + 4 * i;
    fe* m = c->dl + 4 * i;
    fe neg;
    fe_mul(&m[0], &r[3], &c->det[i]);
    fe_neg(&neg, &r[1]);
    fe_mul(&m[1], &neg, &c->det[i]);
    fe_neg(&neg, &r[2]);
    fe_mul(&m[2], &neg, &c->det[i]);
    fe_mul(&m[3], &r[0], &c->det[i]);
  }
}
typedef struct { const point* offset; const point* gen; fe* leaves; } leaf_ctx;
static void leaf_range(size_t lo, size_t hi, void* p) {  /* lib.rs:72-78 from acc = lo * gen */
  leaf_ctx* c = (leaf_ctx*)p;
  point acc = point_zero(), g = *c->gen;
  for (size_t e = lo; e; e >>= 1) {  /* lo * gen by double-and-add: the group law gives the same point */
    if (e & 1) acc = point_add(&acc, &g);
    g = point_add(&g, &g);
  }
  for (size_t i = lo; i < hi; i++) {
    point q = point_add(c->offset, &acc);
    c->leaves[i] = q.x;
    acc = point_add(&acc, c->gen);
  }
}

/* fftree.rs:465-482 */
static orc_tree* derive_subtree(const fe* f, size_t n_parent, const ratmap* maps, size_t nmaps, int parts) {
  size_t n = n_parent / 2;
  if (n == 0) return NULL;
  fe* fp = fe_alloc(2 * n);
  for (size_t i = 0; i < 2 * n; i++) fp[i] = FE_ZERO;
  /* zip(f'_layers, f_layers): layer k of f' (size n>>k at offset n>>k) takes every other
   * element of layer k of f (size n_parent>>k at offset n_parent>>k) */
  for (size_t sz = n, psz = n_parent; sz >= 1; sz >>= 1, psz >>= 1)
    for (size_t i = 0; i < sz; i++) fp[sz + i] = f[psz + 2 * i];
  size_t sub_maps = nmaps ? nmaps - 1 : 0; /* split_last */
  return from_tree(fp, n, maps, sub_maps, parts);
}

/* fftree.rs:318-463.  Takes ownership of f. */
static orc_tree* from_tree(fe* f, size_t n, const ratmap* maps, size_t nmaps, int parts) {
  orc_tree* t = (orc_tree*)calloc(1, sizeof(orc_tree));
  t->sub = derive_subtree(f, n, maps, nmaps, parts);
  t->n = n;
  t->f = f;
  t->nmaps = nmaps;
  t->maps = (ratmap*)malloc((nmaps ? nmaps : 1) * sizeof(ratmap));
  for (size_t i = 0; i < nmaps; i++) t->maps[i] = ratmap_clone(&maps[i]);
  uint64_t nn = n / 2, nnnn = n / 4;
  const fe* s = f + n; /* f_layers[0] */

  /* <X^(n/2) on S>, <X^(n/4) on S> and inverses, fftree.rs:328-333 */
  fe* xnnnn_s = fe_alloc(n);
  fe* xnnnn_s_inv = fe_alloc(n);
  t->xnn_s = fe_alloc(n);
  t->xnn_s_inv = fe_alloc(n);
  {
    pow_ctx pc = {s, xnnnn_s, t->xnn_s, nnnn, nn};
    parallel_for(n, g_build_threads, pow_range, &pc);
  }
  memcpy(xnnnn_s_inv, xnnnn_s, n * sizeof(fe));
  batch_inversion_fast(xnnnn_s_inv, n);
  memcpy(t->xnn_s_inv, t->xnn_s, n * sizeof(fe));
  batch_inversion_fast(t->xnn_s_inv, n);

  /* matrices, Lemma 3.2, fftree.rs:341-363 */
  t->rmat = fe_alloc(4 * n);
  t->dmat = fe_alloc(4 * n);
  for (size_t i = 0; i < n; i++) {
    fe id[4] = {FE_ONE, FE_ZERO, FE_ZERO, FE_ONE};
    memcpy(t->rmat + 4 * i, id, sizeof id);
    memcpy(t->dmat + 4 * i, id, sizeof id);
  }
  for (size_t k = 0; k < nmaps && (n >> (k + 1)) >= 1; k++) {
    const fe* l = f + (n >> k);  /* f layer k, size n>>k */
    size_t d = (n >> k) / 2;
    if (d == 1) continue;
    fe* rl = t->rmat + 4 * d;    /* matrix layer k, size d, at offset d */
    fe* dl = t->dmat + 4 * d;
    fe* det = fe_alloc(d);
    mat_ctx mc = {l, d, &maps[k], rl, dl, det};
    parallel_for(d, g_build_threads, mat_r_range, &mc);
    batch_inversion_fast(det, d);
    parallel_for(d, g_build_threads, mat_d_range, &mc);
    free(det);
  }

  if (parts == 1 || n == 1) {
    free(xnnnn_s);
    free(xnnnn_s_inv);
    return t;
  }

  size_t h = n / 2;
  t->nz = h;
  t->nzz = n;
  t->z0_s1 = fe_alloc(h);
  t->z1_s0 = fe_alloc(h);
  t->z0z0 = fe_alloc(n);
  t->z1z1 = fe_alloc(n);
  fe *s0 = fe_alloc(h), *s1 = fe_alloc(h);
  for (size_t i = 0; i < h; i++) {
    s0[i] = s[2 * i];
    s1[i] = s[2 * i + 1];
  }
  /* <Z_0 on S_1>, <Z_1 on S_0>, fftree.rs:384-405 */
  if (n > 2) {
    const orc_tree* st = t->sub;
    fe *a = fe_alloc(h), *b = fe_alloc(h), *ea = fe_alloc(h), *eb = fe_alloc(h);
    for (size_t i = 0; i < h / 2; i++) {
      a[2 * i] = FE_ZERO;
      a[2 * i + 1] = st->z0_s1[i];
      b[2 * i] = st->z1_s0[i];
      b[2 * i + 1] = FE_ZERO;
    }
    extend_impl(t, a, h, 1, ea, 0);
    extend_impl(t, b, h, 1, eb, 0);
    for (size_t i = 0; i < h; i++) fe_mul(&t->z0_s1[i], &ea[i], &eb[i]);
    fe* z1_s = fe_alloc(n);
    vanish_impl(t, s1, h, z1_s);
    for (size_t i = 0; i < h; i++) t->z1_s0[i] = z1_s[2 * i];
    free(a); free(b); free(ea); free(eb); free(z1_s);
  } else {
    fe_sub(&t->z0_s1[0], &s1[0], &s0[0]);
    fe_sub(&t->z1_s0[0], &s0[0], &s1[0]);
  }
  t->z0_inv_s1 = fe_alloc(h);
  t->z1_inv_s0 = fe_alloc(h);
  memcpy(t->z0_inv_s1, t->z0_s1, h * sizeof(fe));
  memcpy(t->z1_inv_s0, t->z1_s0, h * sizeof(fe));
  batch_inversion_fast(t->z0_inv_s1, h);
  batch_inversion_fast(t->z1_inv_s0, h);

  /* <Z_0^2 mod X^(n/2) on S>, <Z_1^2 mod X^(n/2) on S>, fftree.rs:417-460 */
  if (n > 2) {
    const orc_tree* st = t->sub;
    fe* sq_s0 = fe_alloc(h);
    for (size_t i = 0; i < h; i++) fe_mul(&sq_s0[i], &st->z0z0[i], &st->z1z1[i]);
    fe* r_s0 = fe_alloc(h);
    modular_reduce_impl(st, sq_s0, st->xnn_s, st->z0z0, h, r_s0); /* z0z0_rem_xnnnn_s0 */
    fe* r_s1 = fe_alloc(h);
    extend_impl(t, r_s0, h, 1, r_s1, 0);
    fe* z0z0_rem_xnnnn_s = fe_alloc(n);
    for (size_t i = 0; i < h; i++) {
      z0z0_rem_xnnnn_s[2 * i] = r_s0[i];
      z0z0_rem_xnnnn_s[2 * i + 1] = r_s1[i];
    }
    fe* q = fe_alloc(n); /* ((Z_0 - X^(n/2))^2 - z0z0_rem_xnnnn) / X^(n/4) on S */
    for (size_t i = 0; i < n; i++) {
      fe z0 = (i & 1) ? t->z0_s1[i / 2] : FE_ZERO, d;
      fe_sub(&d, &z0, &t->xnn_s[i]);
      fe_sqr(&d, &d);
      fe_sub(&d, &d, &z0z0_rem_xnnnn_s[i]);
      fe_mul(&q[i], &d, &xnnnn_s_inv[i]);
    }
    fe* qr = fe_alloc(n);
    modular_reduce_impl(t, q, xnnnn_s, z0z0_rem_xnnnn_s, n, qr);
    for (size_t i = 0; i < n; i++) {
      fe m;
      fe_mul(&m, &xnnnn_s[i], &qr[i]);
      fe_add(&t->z0z0[i], &z0z0_rem_xnnnn_s[i], &m);
    }
    fe* z1z1 = fe_alloc(n);
    for (size_t i = 0; i < n; i++) {
      fe z1 = (i & 1) ? FE_ZERO : t->z1_s0[i / 2], d;
      fe_sub(&d, &z1, &t->xnn_s[i]);
      fe_sqr(&z1z1[i], &d);
    }
    modular_reduce_impl(t, z1z1, t->xnn_s, t->z0z0, n, t->z1z1);
    free(sq_s0); free(r_s0); free(r_s1); free(z0z0_rem_xnnnn_s); free(q); free(qr); free(z1z1);
  } else {
    fe a, b;
    fe_sqr(&a, &s0[0]);
    fe_sqr(&b, &s1[0]);
    t->z0z0[0] = t->z0z0[1] = a;
    t->z1z1[0] = t->z1z1[1] = b;
  }
  free(s0); free(s1); free(xnnnn_s); free(xnnnn_s_inv);
  return t;
}

/* FFTree::new, fftree.rs:42-70 */
static orc_tree* fftree_new(const fe* leaves, size_t n, const ratmap* maps, size_t nmaps, int parts) {
  if (!is_pow2(n) || ilog2(n) != nmaps) return NULL;
  fe* f = fe_alloc(2 * n);
  for (size_t i = 0; i < n; i++) f[i] = FE_ZERO;
  memcpy(f + n, leaves, n * sizeof(fe));
  for (size_t k = 0; k < nmaps; k++) {
    const fe* prev = f + (n >> k);
    fe* layer = f + (n >> (k + 1));
    size_t sz = n >> (k + 1);
    /* rational_map.map(prev[i]) = num(x) * den(x)^-1: denominators batch-inverted (same values) */
    fe* den = fe_alloc(sz);
    for (size_t i = 0; i < sz; i++) poly_eval(&den[i], maps[k].den, maps[k].nden, &prev[i]);
    batch_inversion_fast(den, sz);
    for (size_t i = 0; i < sz; i++) {
      fe nu;
      poly_eval(&nu, maps[k].num, maps[k].nnum, &prev[i]);
      fe_mul(&layer[i], &nu, &den[i]);
    }
    free(den);
  }
  return from_tree(f, n, maps, nmaps, parts);
}

/* Fp::build_fftree, reference src/lib.rs:39-85.  The decimal literals of lib.rs:45-59 are
 * given in hex; tests/test_oracle.py::test_curve_constants checks them against the decimals. */
orc_tree* orc_build_fftree(size_t n, int parts) {
  if (!is_pow2(n)) return NULL;
  unsigned log_n = ilog2(n);
  const unsigned two_adicity_of_generator = 36;
  if (log_n >= two_adicity_of_generator) return NULL;
  fe a = fe_from_hex("44eae664a07c69e1c7d7821cacf2a3ccca446568bd32b2a48166309c5c4297e5");
  fe bb = fe_from_hex("649cd342698de65c9bc86f1ece3beb99197d6715a53bdb5609cc937a16154ca8");
  curve c;
  if (!curve_new_odd(&c, &a, &bb)) return NULL;
  point offset, gen;
  offset.inf = gen.inf = 0;
  offset.c = gen.c = c;
  offset.x = fe_from_hex("e9850041b13ea03fadc4bee2afd2959604bf64c290bf3fc15165f15163fd5431");
  offset.y = fe_from_hex("110b996c1374482d0a6b9055a21dc8af9a098495b902b3663322f53ee416d65f");
  gen.x = fe_from_hex("5b4b3e43cd5d95fba244389bb8655539cf8d527f331697e2e93ea60ef50ad5c4");
  gen.y = fe_from_hex("a30fcedca51e68850478e0905816b86d88b79d7b549f4a340016e31de71ded06");
  for (unsigned i = 0; i < two_adicity_of_generator - log_n; i++) gen = point_add(&gen, &gen);

  fe* leaves = fe_alloc(n);
  {
    leaf_ctx lc = {&offset, &gen, leaves};
    parallel_for(n, g_build_threads, leaf_range, &lc);
  }
  /* find_isogeny_chain, ec.rs:177-189 */
  int k = two_adicity(gen);
  if (k != (int)log_n) {
    free(leaves);
    return NULL;
  }
  ratmap* maps = (ratmap*)malloc((k ? k : 1) * sizeof(ratmap));
  point g = gen;
  for (int i = 0; i < k; i++) {
    isogeny iso;
    if (!good_isogeny(&g.c, &iso)) return NULL;
    point gp = isogeny_map(&iso, &g);
    if (two_adicity(g) != two_adicity(gp) + 1) return NULL; /* assert_eq at ec.rs:184 */
    maps[i] = iso.r;
    ratmap_free(&iso.h);
    g = gp;
  }
  orc_tree* t = fftree_new(leaves, n, maps, (size_t)k, parts);
  for (int i = 0; i < k; i++) ratmap_free(&maps[i]);
  free(maps);
  free(leaves);
  return t;
}

/* ------------------------------------------------------------------------- */
/* CanonicalSerialize / CanonicalDeserialize (fftree.rs:510-660) with          */
/* ark-serialize 0.4 conventions: u64 LE lengths, Fp = 32-byte LE canonical    */
/* integer, fixed arrays unprefixed, bool = 1 byte.                            */
/* ------------------------------------------------------------------------- */
typedef struct { uint8_t* p; size_t cap, len; int count_only; } wr;
static void w_bytes(wr* w, const void* src, size_t n) {
  if (!w->count_only && w->len + n <= w->cap) memcpy(w->p + w->len, src, n);
  w->len += n;
}
static void w_u64(wr* w, uint64_t v) {
  uint8_t b[8];
  for (int i = 0; i < 8; i++) b[i] = (uint8_t)(v >> (8 * i));
  w_bytes(w, b, 8);
}
static void w_fe(wr* w, const fe* x) {
  uint8_t b[32];
  if (!w->count_only) orc_fe_to_canonical(x, b);
  w_bytes(w, b, 32);
}
static void w_vec(wr* w, const fe* v, size_t n) {
  w_u64(w, n);
  for (size_t i = 0; i < n; i++) w_fe(w, &v[i]);
}
static void w_mats(wr* w, const fe* v, size_t nmat) {
  w_u64(w, nmat);
  for (size_t i = 0; i < 4 * nmat; i++) w_fe(w, &v[i]);
}
static void w_tree(wr* w, const orc_tree* t, int compressed) {
  w_vec(w, t->f, 2 * t->n);
  w_mats(w, t->rmat, t->n);
  w_mats(w, t->dmat, t->n);
  w_u64(w, t->nmaps);
  for (size_t i = 0; i < t->nmaps; i++) {
    w_vec(w, t->maps[i].num, t->maps[i].nnum);
    w_vec(w, t->maps[i].den, t->maps[i].nden);
  }
  w_vec(w, t->xnn_s, t->n);
  w_vec(w, t->z0_s1, t->nz);
  w_vec(w, t->z1_s0, t->nz);
  if (!compressed) {
    w_vec(w, t->xnn_s_inv, t->n);
    w_vec(w, t->z0_inv_s1, t->nz);
    w_vec(w, t->z1_inv_s0, t->nz);
  }
  w_vec(w, t->z0z0, t->nzz);
  w_vec(w, t->z1z1, t->nzz);
  uint8_t has = t->sub != NULL;
  w_bytes(w, &has, 1);
  if (t->sub) w_tree(w, t->sub, compressed);
}
size_t orc_serialized_size(const orc_tree* t, int compressed) {
  wr w = {NULL, 0, 0, 1};
  w_tree(&w, t, compressed);
  return w.len;
}
size_t orc_serialize(const orc_tree* t, int compressed, uint8_t* buf, size_t cap) {
  wr w = {buf, cap, 0, 0};
  w_tree(&w, t, compressed);
  return w.len;
}

typedef struct { const uint8_t* p; size_t len, pos; int err; } rd;
static uint64_t r_u64(rd* r) {
  if (r->pos + 8 > r->len) { r->err = 1; return 0; }
  uint64_t v = 0;
  for (int i = 7; i >= 0; i--) v = (v << 8) | r->p[r->pos + i];
  r->pos += 8;
  return v;
}
static void r_fe(rd* r, fe* x) {
  if (r->pos + 32 > r->len) { r->err = 1; *x = FE_ZERO; return; }
  /* must be < p (ark-ff rejects non-canonical encodings) */
  uint64_t l[4];
  for (int i = 0; i < 4; i++) {
    uint64_t v = 0;
    for (int k = 7; k >= 0; k--) v = (v << 8) | r->p[r->pos + 8 * i + k];
    l[i] = v;
  }
  if (geq_p(l)) r->err = 1;
  orc_fe_from_canonical(r->p + r->pos, x);
  r->pos += 32;
}
static fe* r_vec(rd* r, size_t* n, size_t per) {
  uint64_t cnt = r_u64(r);
  if (r->err || cnt > (r->len - r->pos) / (32 * per)) { r->err = 1; *n = 0; return fe_alloc(0); }
  *n = (size_t)cnt;
  fe* v = fe_alloc(cnt * per);
  for (size_t i = 0; i < cnt * per; i++) r_fe(r, &v[i]);
  return v;
}
static orc_tree* r_tree(rd* r, int compressed) {
  orc_tree* t = (orc_tree*)calloc(1, sizeof(orc_tree));
  size_t nf, nr, nd, m;
  t->f = r_vec(r, &nf, 1);
  t->rmat = r_vec(r, &nr, 4);
  t->dmat = r_vec(r, &nd, 4);
  t->n = nf / 2;
  uint64_t nm = r_u64(r);
  if (nm > 64) { r->err = 1; nm = 0; }
  t->nmaps = (size_t)nm;
  t->maps = (ratmap*)calloc(nm ? nm : 1, sizeof(ratmap));
  for (size_t i = 0; i < t->nmaps; i++) {
    t->maps[i].num = r_vec(r, &t->maps[i].nnum, 1);
    t->maps[i].den = r_vec(r, &t->maps[i].nden, 1);
  }
  t->xnn_s = r_vec(r, &m, 1);
  if (m != t->n) r->err = 1;
  t->z0_s1 = r_vec(r, &t->nz, 1);
  t->z1_s0 = r_vec(r, &m, 1);
  if (m != t->nz) r->err = 1;
  if (compressed) { /* fftree.rs:621-628 */
    t->xnn_s_inv = fe_alloc(t->n);
    t->z0_inv_s1 = fe_alloc(t->nz);
    t->z1_inv_s0 = fe_alloc(t->nz);
    if (!r->err) {
      memcpy(t->xnn_s_inv, t->xnn_s, t->n * sizeof(fe));
      memcpy(t->z0_inv_s1, t->z0_s1, t->nz * sizeof(fe));
      memcpy(t->z1_inv_s0, t->z1_s0, t->nz * sizeof(fe));
      batch_inversion_fast(t->xnn_s_inv, t->n);
      batch_inversion_fast(t->z0_inv_s1, t->nz);
      batch_inversion_fast(t->z1_inv_s0, t->nz);
    }
  } else {
    t->xnn_s_inv = r_vec(r, &m, 1);
    if (m != t->n) r->err = 1;
    t->z0_inv_s1 = r_vec(r, &m, 1);
    if (m != t->nz) r->err = 1;
    t->z1_inv_s0 = r_vec(r, &m, 1);
    if (m != t->nz) r->err = 1;
  }
  t->z0z0 = r_vec(r, &t->nzz, 1);
  t->z1z1 = r_vec(r, &m, 1);
  if (m != t->nzz) r->err = 1;
  if (nr != t->n || nd != t->n || nf != 2 * t->n) r->err = 1;
  if (r->pos + 1 > r->len) r->err = 1;
  if (!r->err) {
    uint8_t has = r->p[r->pos++];
    if (has > 1) r->err = 1;
    if (has == 1 && !r->err) t->sub = r_tree(r, compressed);
  }
  return t;
}
orc_tree* orc_deserialize(const uint8_t* buf, size_t len, int compressed) {
  rd r = {buf, len, 0, 0};
  orc_tree* t = r_tree(&r, compressed);
  if (r.err) {
    orc_tree_free(t);
    return NULL;
  }
  return t;
}
