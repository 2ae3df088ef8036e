// Short-Weierstrass (a = 0) group law in XYZZ coordinates, generic over the coordinate
// field F (Fp for G1: y^2 = x^3 + 4; Fp2 for G2: y^2 = x^3 + 4(1+u)).
//
// Replaces the reference's affine law with one field inversion per operation
// (impl_affine_add!, src/building_block/curves/macros.rs:35-163).  Every exceptional case
// of that law is reproduced: inf+inf, inf+Q, P+inf (:44-52), P+(-P) -> inf (:53-56),
// P+P -> tangent (:57-108) incl. y = 0 -> inf (:61-63).  Results are converted back to the
// reference's canonical affine form by xyzz_to_affine (one inversion per MSM, not per add).
//
// XYZZ: x = X/ZZ, y = Y/ZZZ with ZZ^3 = ZZZ^2; infinity <=> ZZ = 0.
// Affine points in memory use (0, 0) for AtInfinity (not on either curve since b != 0).
// Costs in field multiplications (M) and squarings (S):
//   madd  (XYZZ += affine)  8M + 2S     add (XYZZ += XYZZ)  12M + 2S
//   dbl   (XYZZ)            6M + 3S     mdbl (2 * affine)    3M + 3S
#pragma once
#include "fp.cuh"
#include "fp2.cuh"

namespace zk {

template <class F> struct Affine { F x, y; };
template <class F> struct XYZZ { F x, y, zz, zzz; };

template <class F> ZK_HD bool is_inf(const Affine<F>& p) { return fis_zero(p.x) && fis_zero(p.y); }
template <class F> ZK_HD bool is_inf(const XYZZ<F>& p) { return fis_zero(p.zz); }
template <class F> ZK_HD void set_inf(XYZZ<F>& p) { fset_zero(p.x); fset_zero(p.y); fset_zero(p.zz); fset_zero(p.zzz); }
template <class F> ZK_HD void set_inf(Affine<F>& p) { fset_zero(p.x); fset_zero(p.y); }

template <class F> ZK_HD void from_affine(XYZZ<F>& r, const Affine<F>& p) {
  if (is_inf(p)) { set_inf(r); return; }
  r.x = p.x; r.y = p.y; fset_one(r.zz); fset_one(r.zzz);
}

// r = 2 * p, p affine and not infinity
template <class F> ZK_HD void xyzz_mdbl(XYZZ<F>& r, const Affine<F>& p) {
  F u, v, w, s, m, t;
  fdbl(u, p.y);
  fsqr(v, u);
  fmul(w, u, v);
  fmul(s, p.x, v);
  fsqr(t, p.x);
  fdbl(m, t); fadd(m, m, t);
  fsqr(t, m);
  fsub(t, t, s); fsub(r.x, t, s);
  fsub(t, s, r.x);
  fmul(t, m, t);
  fmul(u, w, p.y);
  fsub(r.y, t, u);
  r.zz = v;
  r.zzz = w;  // y = 0  =>  v = w = 0  =>  infinity (macros.rs:61-63)
}

// p = 2 * p
template <class F> ZK_HD void xyzz_dbl(XYZZ<F>& p) {
  if (is_inf(p)) return;
  F u, v, w, s, m, t;
  fdbl(u, p.y);
  fsqr(v, u);
  fmul(w, u, v);
  fmul(s, p.x, v);
  fsqr(t, p.x);
  fdbl(m, t); fadd(m, m, t);
  fsqr(t, m);
  fsub(t, t, s); fsub(p.x, t, s);
  fsub(t, s, p.x);
  fmul(t, m, t);
  fmul(u, w, p.y);
  fsub(p.y, t, u);
  fmul(p.zz, v, p.zz);
  fmul(p.zzz, w, p.zzz);
}

// acc += q (affine), complete
template <class F> ZK_HD void xyzz_madd(XYZZ<F>& acc, const Affine<F>& q) {
  if (is_inf(q)) return;
  if (is_inf(acc)) { acc.x = q.x; acc.y = q.y; fset_one(acc.zz); fset_one(acc.zzz); return; }
  F p, r, t, pp, ppp, qq;
  fmul(t, q.x, acc.zz);   fsub(p, t, acc.x);     // P = U2 - X1
  fmul(t, q.y, acc.zzz);  fsub(r, t, acc.y);     // R = S2 - Y1
  if (fis_zero(p)) {
    if (fis_zero(r)) xyzz_mdbl(acc, q);          // same point: tangent (macros.rs:57-108)
    else set_inf(acc);                           // opposite points (macros.rs:53-56)
    return;
  }
  fsqr(pp, p);
  fmul(ppp, p, pp);
  fmul(qq, acc.x, pp);
  fsqr(t, r);
  fsub(t, t, ppp); fsub(t, t, qq); fsub(acc.x, t, qq);   // X3 = R^2 - PPP - 2Q
  fsub(t, qq, acc.x);
  fmul(t, r, t);
  fmul(qq, acc.y, ppp);
  fsub(acc.y, t, qq);                                     // Y3 = R(Q - X3) - Y1 PPP
  fmul(acc.zz, acc.zz, pp);
  fmul(acc.zzz, acc.zzz, ppp);
}

// acc += q (XYZZ), complete
template <class F> ZK_HD void xyzz_add(XYZZ<F>& acc, const XYZZ<F>& q) {
  if (is_inf(q)) return;
  if (is_inf(acc)) { acc = q; return; }
  F u1, s1, p, r, t, pp, ppp, qq;
  fmul(u1, acc.x, q.zz);
  fmul(t, q.x, acc.zz);    fsub(p, t, u1);       // P = U2 - U1
  fmul(s1, acc.y, q.zzz);
  fmul(t, q.y, acc.zzz);   fsub(r, t, s1);       // R = S2 - S1
  if (fis_zero(p)) {
    if (fis_zero(r)) xyzz_dbl(acc);
    else set_inf(acc);
    return;
  }
  fsqr(pp, p);
  fmul(ppp, p, pp);
  fmul(qq, u1, pp);
  fsqr(t, r);
  fsub(t, t, ppp); fsub(t, t, qq); fsub(acc.x, t, qq);
  fsub(t, qq, acc.x);
  fmul(t, r, t);
  fmul(qq, s1, ppp);
  fsub(acc.y, t, qq);
  fmul(t, acc.zz, q.zz);    fmul(acc.zz, t, pp);
  fmul(t, acc.zzz, q.zzz);  fmul(acc.zzz, t, ppp);
}

// ---- latency-oriented variants (independent field products issued as lockstep pairs, fmul2).  Used by the
// kernels after the accumulation, which run a handful of warps and are bound by the length of one
// thread's dependency chain, not by multiplier throughput.  Same formulas, same results.
template <class F> ZK_HD void xyzz_dbl_ilp(XYZZ<F>& p) {
  if (sizeof(F) > 48) { xyzz_dbl(p); return; }   // Fq2: two products in lockstep do not fit the register file (measured 2x slower)
  if (is_inf(p)) return;
  F u, v, w, s, m, t, t2;
  fdbl(u, p.y);
  fmul2(v, u, u, t, p.x, p.x);            // V = U^2 | X^2
  fdbl(m, t); fadd(m, m, t);              // M = 3 X^2
  fmul2(w, u, v, s, p.x, v);              // W = U V | S = X V
  fsqr(t, m);
  fsub(t, t, s); fsub(p.x, t, s);         // X3 = M^2 - 2S
  fsub(t, s, p.x);
  fmul2(t, m, t, u, w, p.y);              // M (S - X3) | W Y
  fsub(p.y, t, u);
  fmul2(p.zz, v, p.zz, p.zzz, w, p.zzz);
  (void)t2;
}

template <class F> ZK_HD void xyzz_add_ilp(XYZZ<F>& acc, const XYZZ<F>& q) {
  if (sizeof(F) > 48) { xyzz_add(acc, q); return; }
  if (is_inf(q)) return;
  if (is_inf(acc)) { acc = q; return; }
  F u1, s1, p, r, t, t2, pp, ppp, qq;
  fmul2(u1, acc.x, q.zz, t, q.x, acc.zz);      fsub(p, t, u1);     // P = U2 - U1
  fmul2(s1, acc.y, q.zzz, t, q.y, acc.zzz);    fsub(r, t, s1);     // R = S2 - S1
  if (fis_zero(p)) {
    if (fis_zero(r)) xyzz_dbl_ilp(acc);
    else set_inf(acc);
    return;
  }
  fmul2(pp, p, p, t, r, r);                    // PP | R^2
  fmul2(ppp, p, pp, qq, u1, pp);               // PPP | Q
  fsub(t, t, ppp); fsub(t, t, qq); fsub(acc.x, t, qq);
  fsub(t, qq, acc.x);
  fmul2(t, r, t, qq, s1, ppp);                 // R (Q - X3) | S1 PPP
  fsub(acc.y, t, qq);
  fmul2(t, acc.zz, q.zz, t2, acc.zzz, q.zzz);
  fmul2(acc.zz, t, pp, acc.zzz, t2, ppp);
}

// canonical affine form; one inversion: t = 1/ZZZ, 1/ZZ = (ZZ t)^2
template <class F> ZK_HD void xyzz_to_affine(Affine<F>& r, const XYZZ<F>& p) {
  if (is_inf(p)) { set_inf(r); return; }
  F t, z;
  finv(t, p.zzz);
  fmul(z, p.zz, t);
  fsqr(z, z);
  fmul(r.x, p.x, z);
  fmul(r.y, p.y, t);
}

template <class F> ZK_HD void affine_cneg(Affine<F>& p, bool neg) { fcneg(p.y, p.y, neg); }

}  // namespace zk
