// Fq2 = Fq[u]/(u^2+1) on top of fp.cuh (reference: curves/bls12_381/fq2.rs:15-151).
// Element = c0 + c1*u; the reference's struct order is (u1, u0) = (c1, c0).
// The reference multiplies schoolbook with 4 Fq products (fq2.rs:139-144); the value is the
// same with Karatsuba's 3, which is what is used here.
#pragma once
#include "fp.cuh"

namespace zk {

struct Fp2 {
  Fp c0, c1;
};

ZK_HD void fset_zero(Fp2& r) { fset_zero(r.c0); fset_zero(r.c1); }
ZK_HD void fset_one(Fp2& r) { fset_one(r.c0); fset_zero(r.c1); }
ZK_HD bool fis_zero(const Fp2& a) { return fis_zero(a.c0) && fis_zero(a.c1); }
ZK_HD bool feq(const Fp2& a, const Fp2& b) { return feq(a.c0, b.c0) && feq(a.c1, b.c1); }
ZK_HD void fadd(Fp2& r, const Fp2& a, const Fp2& b) { fadd(r.c0, a.c0, b.c0); fadd(r.c1, a.c1, b.c1); }
ZK_HD void fsub(Fp2& r, const Fp2& a, const Fp2& b) { fsub(r.c0, a.c0, b.c0); fsub(r.c1, a.c1, b.c1); }
ZK_HD void fdbl(Fp2& r, const Fp2& a) { fdbl(r.c0, a.c0); fdbl(r.c1, a.c1); }
ZK_HD void fneg(Fp2& r, const Fp2& a) { fneg(r.c0, a.c0); fneg(r.c1, a.c1); }
ZK_HD void fcneg(Fp2& r, const Fp2& a, bool neg) { fcneg(r.c0, a.c0, neg); fcneg(r.c1, a.c1, neg); }

ZK_HD void fmul(Fp2& r, const Fp2& a, const Fp2& b) {
  Fp t0, t1, sa, sb, m;
  fmul(t0, a.c0, b.c0);
  fmul(t1, a.c1, b.c1);
  fadd(sa, a.c0, a.c1);
  fadd(sb, b.c0, b.c1);
  fmul(m, sa, sb);
  fsub(m, m, t0);
  fsub(r.c1, m, t1);
  fsub(r.c0, t0, t1);
}

// pair of independent Fq2 products (no lockstep variant: the Fq2 product already has 3 independent Fq products)
ZK_HD void fmul2(Fp2& r1, const Fp2& a1, const Fp2& b1, Fp2& r2, const Fp2& a2, const Fp2& b2) {
  Fp2 t1, t2;
  fmul(t1, a1, b1);
  fmul(t2, a2, b2);
  r1 = t1;
  r2 = t2;
}

ZK_HD void fsqr(Fp2& r, const Fp2& a) {
  Fp s, d, m;
  fadd(s, a.c0, a.c1);
  fsub(d, a.c0, a.c1);
  fmul(m, a.c0, a.c1);
  fmul(r.c0, s, d);
  fdbl(r.c1, m);
}

// fq2.rs:26-32: (c0 - c1 u) / (c0^2 + c1^2)
ZK_HD void finv(Fp2& r, const Fp2& a) {
  Fp n0, n1, t;
  fsqr(n0, a.c0);
  fsqr(n1, a.c1);
  fadd(n0, n0, n1);
  finv(t, n0);
  fmul(r.c0, a.c0, t);
  fmul(n1, a.c1, t);
  fneg(r.c1, n1);
}

}  // namespace zk
