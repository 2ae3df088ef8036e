/* T0 oracle in C: a CPU restatement of the reference's MSM algorithm, operation for operation.
 *
 * TEST INFRASTRUCTURE ONLY.  Used by tests/ (as a checker) and by bench.py's cpu_baseline /
 * --impl reference legs (as the reported CPU baseline).  Never linked into the product.
 *
 * The reference is Rust on num-bigint and cannot be built here (no cargo/rustc, un-vendored git
 * dependency), so this file restates its algorithm on fixed-size integers:
 *   field mul   = full product then `% q` by long division      (prime_field_elem.rs:302-308)
 *   field inv   = extended Euclid on signed integers, one division per step (:379-432)
 *   add/sub     = one conditional correction                     (:278-300)
 *   group add   = affine chord / tangent with one inversion each (curves/macros.rs:35-163)
 *   scalar mul  = LSB-first double-and-add on the raw integer    (curves/macros.rs:2-32)
 *   MSM         = serial sum of per-term scalar multiplications  (field/polynomial.rs:272-293)
 * Pinned by the reference's golden vectors (tests/test_oracle_c.py) and by agreement with the
 * Python T0 oracle (oracle/zkt_oracle.py) on random inputs.
 *
 * Threads: the reference is single-threaded.  With threads > 1 the term range is split evenly,
 * each thread runs the reference loop on its slice, and the partial sums are added in order.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef uint64_t u64;
typedef unsigned __int128 u128;

#define NL 6  /* 64-bit limbs of an Fq element */
#define WL 14 /* working width for the Euclid sequences and products */

static const u64 Q[NL] = {0xb9feffffffffaaabull, 0x1eabfffeb153ffffull, 0x6730d2a0f6b0f624ull,
                          0x64774b84f38512bfull, 0x4b1ba7b6434bacd7ull, 0x1a0111ea397fe69aull};

/* ------------------------------------------------------------------ unsigned multi-limb helpers */
static int bn_len(const u64* a, int n) { while (n > 0 && a[n - 1] == 0) n--; return n; }
static int bn_cmp(const u64* a, const u64* b, int n) {
  for (int i = n - 1; i >= 0; i--) if (a[i] != b[i]) return a[i] < b[i] ? -1 : 1;
  return 0;
}
static int bn_is_zero(const u64* a, int n) { return bn_len(a, n) == 0; }
static u64 bn_add(u64* r, const u64* a, const u64* b, int n) {
  u64 c = 0;
  for (int i = 0; i < n; i++) { u128 t = (u128)a[i] + b[i] + c; r[i] = (u64)t; c = (u64)(t >> 64); }
  return c;
}
static u64 bn_sub(u64* r, const u64* a, const u64* b, int n) {
  u64 bw = 0;
  for (int i = 0; i < n; i++) { u128 t = (u128)a[i] - b[i] - bw; r[i] = (u64)t; bw = (u64)(t >> 64) & 1; }
  return bw;
}
/* r[na+nb] = a * b */
static void bn_mul(u64* r, const u64* a, int na, const u64* b, int nb) {
  memset(r, 0, sizeof(u64) * (na + nb));
  for (int i = 0; i < na; i++) {
    u64 c = 0;
    for (int j = 0; j < nb; j++) { u128 t = (u128)a[i] * b[j] + r[i + j] + c; r[i + j] = (u64)t; c = (u64)(t >> 64); }
    r[i + nb] = c;
  }
}
/* Knuth algorithm D: q[nn] = num / den, r[nd] = num % den (den != 0); limbs above the true length are zeroed */
static void bn_divmod(u64* q, u64* r, const u64* num, int nn, const u64* den, int nd) {
  int ln = bn_len(num, nn), ld = bn_len(den, nd);
  if (q) memset(q, 0, sizeof(u64) * nn);
  memset(r, 0, sizeof(u64) * nd);
  if (ln < ld || (ln == ld && bn_cmp(num, den, ld) < 0)) { memcpy(r, num, sizeof(u64) * ln); return; }
  if (ld == 1) {
    u64 rem = 0;
    for (int i = ln - 1; i >= 0; i--) { u128 t = ((u128)rem << 64) | num[i]; u64 qq = (u64)(t / den[0]); rem = (u64)(t % den[0]); if (q) q[i] = qq; }
    r[0] = rem;
    return;
  }
  u64 un[2 * WL + 2], vn[WL + 1];
  int s = __builtin_clzll(den[ld - 1]);
  for (int i = ld - 1; i > 0; i--) vn[i] = s ? (den[i] << s) | (den[i - 1] >> (64 - s)) : den[i];
  vn[0] = den[0] << s;
  un[ln] = s ? num[ln - 1] >> (64 - s) : 0;
  for (int i = ln - 1; i > 0; i--) un[i] = s ? (num[i] << s) | (num[i - 1] >> (64 - s)) : num[i];
  un[0] = num[0] << s;
  for (int j = ln - ld; j >= 0; j--) {
    u128 numer = ((u128)un[j + ld] << 64) | un[j + ld - 1];
    u128 qhat = numer / vn[ld - 1], rhat = numer % vn[ld - 1];
    while ((qhat >> 64) || (u128)(u64)qhat * vn[ld - 2] > ((rhat << 64) | un[j + ld - 2])) {
      qhat--;
      rhat += vn[ld - 1];
      if (rhat >> 64) break;
    }
    u64 qh = (u64)qhat, carry = 0, bw = 0;
    for (int i = 0; i < ld; i++) {
      u128 p = (u128)qh * vn[i] + carry;
      carry = (u64)(p >> 64);
      u128 t = (u128)un[i + j] - (u64)p - bw;
      un[i + j] = (u64)t;
      bw = (u64)(t >> 64) & 1;
    }
    u128 t = (u128)un[j + ld] - carry - bw;
    un[j + ld] = (u64)t;
    if ((u64)(t >> 64) & 1) { /* qhat was one too large: add the divisor back */
      qh--;
      u64 c = 0;
      for (int i = 0; i < ld; i++) { u128 a = (u128)un[i + j] + vn[i] + c; un[i + j] = (u64)a; c = (u64)(a >> 64); }
      un[j + ld] += c;
    }
    if (q) q[j] = qh;
  }
  for (int i = 0; i < ld; i++) r[i] = s ? (un[i] >> s) | ((i + 1 <= ld ? un[i + 1] : 0) << (64 - s)) : un[i];
}

/* ------------------------------------------------------------------ Fq (PrimeFieldElem over q) */
typedef struct { u64 v[NL]; } fq;

static void fq_add(fq* r, const fq* a, const fq* b) { /* plus, :278-286 */
  u64 t[NL];
  u64 c = bn_add(t, a->v, b->v, NL);
  if (c || bn_cmp(t, Q, NL) >= 0) bn_sub(t, t, Q, NL);
  memcpy(r->v, t, sizeof(t));
}
static void fq_sub(fq* r, const fq* a, const fq* b) { /* minus, :288-300 */
  u64 t[NL];
  if (bn_cmp(a->v, b->v, NL) < 0) { bn_sub(t, b->v, a->v, NL); bn_sub(t, Q, t, NL); }
  else bn_sub(t, a->v, b->v, NL);
  memcpy(r->v, t, sizeof(t));
}
static void fq_mul(fq* r, const fq* a, const fq* b) { /* times, :302-308: e *= rhs; e %= order */
  u64 p[2 * NL], rem[NL];
  bn_mul(p, a->v, NL, b->v, NL);
  bn_divmod(NULL, rem, p, 2 * NL, Q, NL);
  memcpy(r->v, rem, sizeof(rem));
}
static void fq_sq(fq* r, const fq* a) { fq_mul(r, a, a); } /* sq, :330-335 */
static void fq_neg(fq* r, const fq* a) {                   /* negate, :448-457 */
  if (bn_is_zero(a->v, NL)) { *r = *a; return; }
  bn_sub(r->v, Q, a->v, NL);
}
static int fq_is_zero(const fq* a) { return bn_is_zero(a->v, NL); }
static int fq_eq(const fq* a, const fq* b) { return bn_cmp(a->v, b->v, NL) == 0; }

/* signed integer for the Euclid cofactors */
typedef struct { int neg; u64 m[WL]; } sbn;
static void sbn_sub_mul(sbn* r, const sbn* x0, const sbn* x1, const u64* q, int nq) { /* r = x0 - x1*q */
  u64 prod[2 * WL];
  int l1 = bn_len(x1->m, WL), lq = bn_len(q, nq);
  memset(prod, 0, sizeof(prod));
  if (l1 && lq) bn_mul(prod, x1->m, l1, q, lq);
  int pneg = x1->neg; /* sign of x1*q (q >= 0) */
  /* r = x0 + (-(x1*q)) */
  int bneg = !pneg;
  if (bn_is_zero(prod, WL)) { *r = *x0; return; }
  if (x0->neg == bneg) { bn_add(r->m, x0->m, prod, WL); r->neg = bneg; }
  else if (bn_cmp(x0->m, prod, WL) >= 0) { bn_sub(r->m, x0->m, prod, WL); r->neg = x0->neg; }
  else { bn_sub(r->m, prod, x0->m, WL); r->neg = bneg; }
  if (bn_is_zero(r->m, WL)) r->neg = 0;
}

/* safe_inv, prime_field_elem.rs:379-432.  Returns 0 on success, -1 for zero ("Cannot find inverse of zero"). */
static int fq_inv(fq* out, const fq* a) {
  if (fq_is_zero(a)) return -1;
  u64 r0[WL] = {0}, r1[WL] = {0}, qq[WL], r2[WL];
  memcpy(r0, a->v, sizeof(u64) * NL);
  memcpy(r1, Q, sizeof(u64) * NL);
  sbn x0 = {0, {1}}, x1 = {0, {0}}, x2;
  /* the y sequence of the reference does not influence the result; it is kept for cost parity */
  sbn y0 = {0, {0}}, y1 = {0, {1}}, y2;
  while (!bn_is_zero(r1, WL)) {
    bn_divmod(qq, r2, r0, WL, r1, WL);          /* q = r0 / r1 ; r2 = r0 % r1 */
    sbn_sub_mul(&x2, &x0, &x1, qq, WL);
    sbn_sub_mul(&y2, &y0, &y1, qq, WL);
    memcpy(r0, r1, sizeof(r0));
    memcpy(r1, r2, sizeof(r1));
    x0 = x1; y0 = y1; x1 = x2; y1 = y2;
  }
  u64 v[WL];
  memcpy(v, x0.m, sizeof(v));
  if (x0.neg) {             /* while new_v < 0 { new_v += order } */
    u64 qext[WL] = {0};
    memcpy(qext, Q, sizeof(u64) * NL);
    int neg = 1;
    while (neg) {
      if (bn_cmp(v, qext, WL) <= 0) { bn_sub(v, qext, v, WL); neg = 0; }
      else bn_sub(v, v, qext, WL);
    }
  } else {
    u64 qext[WL] = {0}, rem[WL];
    memcpy(qext, Q, sizeof(u64) * NL);
    if (bn_cmp(v, qext, WL) >= 0) { bn_divmod(NULL, rem, v, WL, qext, WL); memcpy(v, rem, sizeof(v)); }
  }
  memcpy(out->v, v, sizeof(u64) * NL);
  return 0;
}

/* ------------------------------------------------------------------ Fq2 (fq2.rs), element = (u1, u0) */
typedef struct { fq u1, u0; } fq2;
static void fq2_add(fq2* r, const fq2* a, const fq2* b) { fq_add(&r->u1, &a->u1, &b->u1); fq_add(&r->u0, &a->u0, &b->u0); }
static void fq2_sub(fq2* r, const fq2* a, const fq2* b) { fq_sub(&r->u1, &a->u1, &b->u1); fq_sub(&r->u0, &a->u0, &b->u0); }
static void fq2_mul(fq2* r, const fq2* a, const fq2* b) { /* fq2.rs:134-151: four products */
  fq t0, t1, t2, t3, o1, o0;
  fq_mul(&t0, &a->u1, &b->u0); fq_mul(&t1, &a->u0, &b->u1); fq_add(&o1, &t0, &t1);
  fq_mul(&t2, &a->u0, &b->u0); fq_mul(&t3, &a->u1, &b->u1); fq_sub(&o0, &t2, &t3);
  r->u1 = o1; r->u0 = o0;
}
static void fq2_sq(fq2* r, const fq2* a) { fq2_mul(r, a, a); } /* fq2.rs:34-36 */
static void fq2_neg(fq2* r, const fq2* a) { fq2 z; memset(&z, 0, sizeof(z)); fq2_sub(r, &z, a); } /* :82-94 */
static int fq2_inv(fq2* r, const fq2* a) { /* fq2.rs:26-32 */
  fq s1, s0, n, f, t;
  fq_mul(&s1, &a->u1, &a->u1); fq_mul(&s0, &a->u0, &a->u0); fq_add(&n, &s1, &s0);
  if (fq_inv(&f, &n)) return -1;
  fq_neg(&t, &a->u1); fq_mul(&r->u1, &t, &f);
  fq_mul(&r->u0, &a->u0, &f);
  return 0;
}
static int fq2_is_zero(const fq2* a) { return fq_is_zero(&a->u0) && fq_is_zero(&a->u1); }
static int fq2_eq(const fq2* a, const fq2* b) { return fq_eq(&a->u0, &b->u0) && fq_eq(&a->u1, &b->u1); }

/* ------------------------------------------------------------------ affine group law, generic by macro */
#define DEFINE_GROUP(P, F, F_add, F_sub, F_mul, F_sq, F_neg, F_inv, F_is_zero, F_eq)                         \
  typedef struct { int inf; F x, y; } P;                                                                      \
  static void P##_add(P* r, const P* p, const P* q) { /* impl_affine_add!, macros.rs:35-163 */               \
    if (p->inf && q->inf) { r->inf = 1; return; }                                                             \
    if (p->inf) { *r = *q; return; }                                                                          \
    if (q->inf) { *r = *p; return; }                                                                          \
    int same_x = F_eq(&p->x, &q->x), same_y = F_eq(&p->y, &q->y);                                             \
    if (same_x && !same_y) { r->inf = 1; return; }                                                            \
    F m, t, u, x3, y3;                                                                                        \
    if (same_x && same_y) {                                                                                   \
      if (F_is_zero(&p->y)) { r->inf = 1; return; }                                                           \
      F_sq(&t, &p->x); F_add(&u, &t, &t); F_add(&u, &u, &t);        /* m1 = 3 x^2 */                          \
      F_add(&t, &p->y, &p->y);                                      /* m2 = 2 y   */                          \
      F_inv(&t, &t); F_mul(&m, &u, &t);                                                                       \
      F_sq(&t, &m); F_add(&u, &p->x, &p->x); F_sub(&x3, &t, &u);    /* m^2 - 2x   */                          \
      F_sub(&t, &p->x, &x3); F_mul(&t, &m, &t); F_sub(&y3, &t, &p->y);                                        \
      r->inf = 0; r->x = x3; r->y = y3; return;                                                               \
    }                                                                                                         \
    F_sub(&t, &q->y, &p->y); F_sub(&u, &q->x, &p->x); F_inv(&u, &u); F_mul(&m, &t, &u);                      \
    F_sq(&t, &m); F_sub(&t, &t, &p->x); F_sub(&x3, &t, &q->x);                                                \
    F_sub(&t, &x3, &p->x); F_mul(&t, &m, &t); F_add(&t, &t, &p->y); F_neg(&y3, &t);                           \
    r->inf = 0; r->x = x3; r->y = y3;                                                                         \
  }                                                                                                           \
  /* impl_scalar_mul_point!, macros.rs:2-32: scan the raw 256-bit integer from the LSB */                     \
  static void P##_mul(P* r, const P* p, const uint32_t k[8]) {                                                \
    P res, pw = *p, t;                                                                                        \
    res.inf = 1;                                                                                              \
    int top = 255;                                                                                            \
    while (top >= 0 && !((k[top >> 5] >> (top & 31)) & 1)) top--;                                             \
    for (int b = 0; b <= top; b++) {                                                                          \
      if ((k[b >> 5] >> (b & 31)) & 1) { P##_add(&t, &res, &pw); res = t; }                                   \
      P##_add(&t, &pw, &pw); pw = t;                                                                          \
    }                                                                                                         \
    *r = res;                                                                                                 \
  }

DEFINE_GROUP(g1p, fq, fq_add, fq_sub, fq_mul, fq_sq, fq_neg, fq_inv, fq_is_zero, fq_eq)
DEFINE_GROUP(g2p, fq2, fq2_add, fq2_sub, fq2_mul, fq2_sq, fq2_neg, fq2_inv, fq2_is_zero, fq2_eq)

/* ------------------------------------------------------------------ flat-limb I/O (the C ABI's layout) */
static void fq_load(fq* r, const uint32_t* w) { for (int i = 0; i < NL; i++) r->v[i] = (u64)w[2 * i] | ((u64)w[2 * i + 1] << 32); }
static void fq_store(uint32_t* w, const fq* a) { for (int i = 0; i < NL; i++) { w[2 * i] = (uint32_t)a->v[i]; w[2 * i + 1] = (uint32_t)(a->v[i] >> 32); } }
static void g1_load(g1p* p, const uint32_t* w, int inf) { p->inf = inf; if (!inf) { fq_load(&p->x, w); fq_load(&p->y, w + 12); } }
static void g1_store(uint32_t* w, int* inf, const g1p* p) { *inf = p->inf; memset(w, 0, 96); if (!p->inf) { fq_store(w, &p->x); fq_store(w + 12, &p->y); } }
/* G2 ABI order x.u0 | x.u1 | y.u0 | y.u1 */
static void g2_load(g2p* p, const uint32_t* w, int inf) {
  p->inf = inf;
  if (!inf) { fq_load(&p->x.u0, w); fq_load(&p->x.u1, w + 12); fq_load(&p->y.u0, w + 24); fq_load(&p->y.u1, w + 36); }
}
static void g2_store(uint32_t* w, int* inf, const g2p* p) {
  *inf = p->inf; memset(w, 0, 192);
  if (!p->inf) { fq_store(w, &p->x.u0); fq_store(w + 12, &p->x.u1); fq_store(w + 24, &p->y.u0); fq_store(w + 36, &p->y.u1); }
}

/* ------------------------------------------------------------------ MSM = eval_with_g{1,2}_hidings */
typedef struct { const uint32_t* xy; const uint8_t* inf; const uint32_t* sc; size_t lo, hi; g1p out1; g2p out2; } job;

static void* g1_worker(void* arg) {
  job* j = (job*)arg;
  g1p sum, p, t, u;
  sum.inf = 1;
  for (size_t i = j->lo; i < j->hi; i++) { /* sum = sum + (&powers[i] * &coeffs[i]), polynomial.rs:277-279 */
    g1_load(&p, j->xy + 24 * i, j->inf ? j->inf[i] : 0);
    g1p_mul(&t, &p, j->sc + 8 * i);
    g1p_add(&u, &sum, &t);
    sum = u;
  }
  j->out1 = sum;
  return NULL;
}
static void* g2_worker(void* arg) {
  job* j = (job*)arg;
  g2p sum, p, t, u;
  sum.inf = 1;
  for (size_t i = j->lo; i < j->hi; i++) {
    g2_load(&p, j->xy + 48 * i, j->inf ? j->inf[i] : 0);
    g2p_mul(&t, &p, j->sc + 8 * i);
    g2p_add(&u, &sum, &t);
    sum = u;
  }
  j->out2 = sum;
  return NULL;
}

static int run_jobs(job* jobs, int threads, void* (*fn)(void*), const uint32_t* xy, const uint8_t* inf, const uint32_t* sc, size_t n) {
  pthread_t th[256];
  if (threads < 1) threads = 1;
  if (threads > 256) threads = 256;
  if ((size_t)threads > n) threads = n ? (int)n : 1;
  for (int t = 0; t < threads; t++) {
    jobs[t].xy = xy; jobs[t].inf = inf; jobs[t].sc = sc;
    jobs[t].lo = n * t / threads; jobs[t].hi = n * (t + 1) / threads;
  }
  if (threads == 1) { fn(&jobs[0]); return 1; }
  for (int t = 0; t < threads; t++) pthread_create(&th[t], NULL, fn, &jobs[t]);
  for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
  return threads;
}

int zkt_g1_msm_ref(const uint32_t* xy, const uint8_t* inf, const uint32_t* scalars, size_t n, int threads, uint32_t out_xy[24], int* out_inf) {
  job jobs[256];
  int used = run_jobs(jobs, threads, g1_worker, xy, inf, scalars, n);
  g1p sum = jobs[0].out1, t;
  for (int i = 1; i < used; i++) { g1p_add(&t, &sum, &jobs[i].out1); sum = t; }
  g1_store(out_xy, out_inf, &sum);
  return 0;
}
int zkt_g2_msm_ref(const uint32_t* xy, const uint8_t* inf, const uint32_t* scalars, size_t n, int threads, uint32_t out_xy[48], int* out_inf) {
  job jobs[256];
  int used = run_jobs(jobs, threads, g2_worker, xy, inf, scalars, n);
  g2p sum = jobs[0].out2, t;
  for (int i = 1; i < used; i++) { g2p_add(&t, &sum, &jobs[i].out2); sum = t; }
  g2_store(out_xy, out_inf, &sum);
  return 0;
}

/* single operations, for the golden-vector tests */
int zkt_g1_add_ref(const uint32_t* p, int pinf, const uint32_t* q, int qinf, uint32_t out[24], int* out_inf) {
  g1p a, b, r; g1_load(&a, p, pinf); g1_load(&b, q, qinf); g1p_add(&r, &a, &b); g1_store(out, out_inf, &r); return 0;
}
int zkt_g1_mul_ref(const uint32_t* p, int pinf, const uint32_t k[8], uint32_t out[24], int* out_inf) {
  g1p a, r; g1_load(&a, p, pinf); g1p_mul(&r, &a, k); g1_store(out, out_inf, &r); return 0;
}
int zkt_g2_add_ref(const uint32_t* p, int pinf, const uint32_t* q, int qinf, uint32_t out[48], int* out_inf) {
  g2p a, b, r; g2_load(&a, p, pinf); g2_load(&b, q, qinf); g2p_add(&r, &a, &b); g2_store(out, out_inf, &r); return 0;
}
int zkt_g2_mul_ref(const uint32_t* p, int pinf, const uint32_t k[8], uint32_t out[48], int* out_inf) {
  g2p a, r; g2_load(&a, p, pinf); g2p_mul(&r, &a, k); g2_store(out, out_inf, &r); return 0;
}
/* op: 0 add 1 sub 2 mul 3 neg 4 inv */
int zkt_fq_op_ref(int op, const uint32_t a[12], const uint32_t b[12], uint32_t out[12]) {
  fq x, y, r; fq_load(&x, a); if (b) fq_load(&y, b);
  int rc = 0;
  switch (op) {
    case 0: fq_add(&r, &x, &y); break;
    case 1: fq_sub(&r, &x, &y); break;
    case 2: fq_mul(&r, &x, &y); break;
    case 3: fq_neg(&r, &x); break;
    case 4: rc = fq_inv(&r, &x); if (rc) memset(&r, 0, sizeof(r)); break;
    default: return -2;
  }
  fq_store(out, &r);
  return rc;
}
