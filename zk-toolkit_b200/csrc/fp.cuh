// Montgomery prime-field arithmetic on 32-bit limbs for sm_100a.
//
// Replaces the reference's BigUint field ops on the MSM path
// (src/building_block/field/prime_field_elem.rs:278-335 plus/minus/times/sq,
// :448-457 negate, :379-436 inv) for the two BLS12-381 fields of params.rs:9,14:
// Fq (381 bit, 12 limbs) and Fr (255 bit, 8 limbs).
//
// Representation: a*R mod p, R = 2^(32 N), little-endian limbs, ALWAYS fully reduced
// to [0, p) so that zero / equality tests are exact limb comparisons.
//
// Multiplication is operand-scanning Montgomery with the running sum split into two
// accumulators whose 64-bit digits sit at even resp. odd limb offsets.  Every limb
// product a_j*b_i is then a 64-bit-aligned add into one of them, i.e. one
// IMAD.WIDE.U32.X on a carry chain; nothing ever has to be split into lo/hi halves.
// Cost per modmul with N = 12: 2*N*N = 288 IMAD.WIDE + N IMAD (the m_i) on the fma pipe
// and ~90 IADD3/LOP3 on the alu pipe.  Limb-product (LP) count used by the roofline
// in DESIGN.md: 2 N^2 + N = 300.
#pragma once
#include "ptx.cuh"
#include "constants.cuh"

namespace zk {

struct FqCfg {
  static constexpr int N = 12;
  static constexpr uint32_t INV = FQ_INV;
  ZK_HD static uint32_t p(int i) { return FQ_P[i]; }
  ZK_HD static uint32_t one(int i) { return FQ_ONE[i]; }
  ZK_HD static uint32_t r2(int i) { return FQ_R2[i]; }
  ZK_HD static uint32_t pm2(int i) { return FQ_PM2[i]; }
  ZK_HD static uint32_t r3(int i) { return FQ_R3[i]; }
};

struct FrCfg {
  static constexpr int N = 8;
  static constexpr uint32_t INV = FR_INV;
  ZK_HD static uint32_t p(int i) { return FR_P[i]; }
  ZK_HD static uint32_t one(int i) { return FR_ONE[i]; }
  ZK_HD static uint32_t r2(int i) { return FR_R2[i]; }
  ZK_HD static uint32_t pm2(int i) { return FR_PM2[i]; }
  ZK_HD static uint32_t r3(int i) { return FR_R3[i]; }
};

template <class Cfg>
struct alignas(16) Mont {
  static constexpr int N = Cfg::N;
  typedef Cfg cfg;
  uint32_t v[Cfg::N];
};

typedef Mont<FqCfg> Fp;
typedef Mont<FrCfg> Fr;

// ---------------------------------------------------------------- small helpers
template <class C> ZK_HD void fset_zero(Mont<C>& r) {
#pragma unroll
  for (int i = 0; i < C::N; i++) r.v[i] = 0;
}
template <class C> ZK_HD void fset_one(Mont<C>& r) {
#pragma unroll
  for (int i = 0; i < C::N; i++) r.v[i] = C::one(i);
}
template <class C> ZK_HD bool fis_zero(const Mont<C>& a) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < C::N; i++) o |= a.v[i];
  return o == 0;
}
template <class C> ZK_HD bool feq(const Mont<C>& a, const Mont<C>& b) {
  uint32_t o = 0;
#pragma unroll
  for (int i = 0; i < C::N; i++) o |= a.v[i] ^ b.v[i];
  return o == 0;
}

// r = (t >= p) ? t - p : t, for t < 2p
template <class C> ZK_HD void freduce_once(Mont<C>& r, const uint32_t* t) {
  uint32_t d[C::N];
  d[0] = ptx::sub_cc(t[0], C::p(0));
#pragma unroll
  for (int i = 1; i < C::N; i++) d[i] = ptx::subc_cc(t[i], C::p(i));
  uint32_t borrow = ptx::subc(0, 0);  // all-ones when t < p
#pragma unroll
  for (int i = 0; i < C::N; i++) r.v[i] = (t[i] & borrow) | (d[i] & ~borrow);
}

// ---------------------------------------------------------------- add / sub / neg
template <class C> ZK_HD void fadd(Mont<C>& r, const Mont<C>& a, const Mont<C>& b) {
  uint32_t t[C::N];
  t[0] = ptx::add_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < C::N - 1; i++) t[i] = ptx::addc_cc(a.v[i], b.v[i]);
  t[C::N - 1] = ptx::addc(a.v[C::N - 1], b.v[C::N - 1]);  // p has >= 1 spare top bit: no carry out
  freduce_once(r, t);
}

template <class C> ZK_HD void fsub(Mont<C>& r, const Mont<C>& a, const Mont<C>& b) {
  uint32_t t[C::N];
  t[0] = ptx::sub_cc(a.v[0], b.v[0]);
#pragma unroll
  for (int i = 1; i < C::N; i++) t[i] = ptx::subc_cc(a.v[i], b.v[i]);
  uint32_t borrow = ptx::subc(0, 0);  // all-ones when a < b: add p back
  r.v[0] = ptx::add_cc(t[0], C::p(0) & borrow);
#pragma unroll
  for (int i = 1; i < C::N - 1; i++) r.v[i] = ptx::addc_cc(t[i], C::p(i) & borrow);
  r.v[C::N - 1] = ptx::addc(t[C::N - 1], C::p(C::N - 1) & borrow);
}

template <class C> ZK_HD void fdbl(Mont<C>& r, const Mont<C>& a) { fadd(r, a, a); }

// -a, with -0 = 0 (prime_field_elem.rs:448-457)
template <class C> ZK_HD void fneg(Mont<C>& r, const Mont<C>& a) {
  uint32_t nz = fis_zero(a) ? 0u : 0xffffffffu;
  uint32_t t[C::N];
  t[0] = ptx::sub_cc(C::p(0), a.v[0]);
#pragma unroll
  for (int i = 1; i < C::N - 1; i++) t[i] = ptx::subc_cc(C::p(i), a.v[i]);
  t[C::N - 1] = ptx::subc(C::p(C::N - 1), a.v[C::N - 1]);
#pragma unroll
  for (int i = 0; i < C::N; i++) r.v[i] = t[i] & nz;
}

// r = neg ? -a : a
template <class C> ZK_HD void fcneg(Mont<C>& r, const Mont<C>& a, bool neg) {
  Mont<C> n;
  fneg(n, a);
  uint32_t m = neg ? 0xffffffffu : 0u;
#pragma unroll
  for (int i = 0; i < C::N; i++) r.v[i] = (n.v[i] & m) | (a.v[i] & ~m);
}

// ---------------------------------------------------------------- Montgomery multiplication
namespace detail {

// acc[0..n) = x[0], x[2], .. , x[n-2] times y as consecutive 64-bit digits (no accumulate)
template <int n> ZK_HD void row_mul(uint32_t* acc, const uint32_t* x, uint32_t y) {
#pragma unroll
  for (int j = 0; j < n; j += 2) {
    acc[j] = ptx::mul_lo(x[j], y);
    acc[j + 1] = ptx::mul_hi(x[j], y);
  }
}

// acc[0..n) += (x[0], x[2], ..) * y on one carry chain; the chain's carry-out stays in CC
template <int n, class X> ZK_HD void row_mad(uint32_t* acc, X x, uint32_t y) {
  acc[0] = ptx::mad_lo_cc(x(0), y, acc[0]);
  acc[1] = ptx::madc_hi_cc(x(0), y, acc[1]);
#pragma unroll
  for (int j = 2; j < n; j += 2) {
    acc[j] = ptx::madc_lo_cc(x(j), y, acc[j]);
    acc[j + 1] = ptx::madc_hi_cc(x(j), y, acc[j + 1]);
  }
}

// acc = (acc >> 64) + (x[0], x[2], ..) * y, consuming the carry already in CC; no carry-out
template <int n, class X> ZK_HD void row_mad_shift(uint32_t* acc, X x, uint32_t y) {
#pragma unroll
  for (int j = 0; j < n - 2; j += 2) {
    acc[j] = ptx::madc_lo_cc(x(j), y, acc[j + 2]);
    acc[j + 1] = ptx::madc_hi_cc(x(j), y, acc[j + 3]);
  }
  acc[n - 2] = ptx::madc_lo_cc(x(n - 2), y, 0);
  acc[n - 1] = ptx::madc_hi(x(n - 2), y, 0);
}

template <class C> struct ModEven { ZK_HD uint32_t operator()(int j) const { return C::p(j); } };
template <class C> struct ModOdd { ZK_HD uint32_t operator()(int j) const { return C::p(j + 1); } };
struct Arr {
  const uint32_t* p;
  ZK_HD uint32_t operator()(int j) const { return p[j]; }
};

// One operand-scanning step: T = (T + a*bi + m*p) / 2^32 in split form.
// On entry (not first) the running sum is  lo + hi*2^32  where `lo` still holds the previous
// step's even digits (its limb 0 is zero, limb 1 is the odd leftover) and `hi` the odd digits.
// The division by 2^32 is the renaming: hi becomes the new even accumulator E, lo>>64 the new odd O.
template <class C> ZK_HD void mont_step(uint32_t* E, uint32_t* O, const uint32_t* a, uint32_t bi, bool first) {
  constexpr int n = C::N;
  if (first) {
    row_mul<n>(O, a + 1, bi);
    row_mul<n>(E, a, bi);
  } else {
    E[0] = ptx::add_cc(E[0], O[1]);
    row_mad_shift<n>(O, Arr{a + 1}, bi);
    row_mad<n>(E, Arr{a}, bi);
    O[n - 1] = ptx::addc(O[n - 1], 0);
  }
  uint32_t m = E[0] * C::INV;
  row_mad<n>(O, ModOdd<C>(), m);       // cannot carry out: total < 2^(32(n+1))
  row_mad<n>(E, ModEven<C>(), m);
  O[n - 1] = ptx::addc(O[n - 1], 0);
}

}  // namespace detail

// Translation units off the hot loop define ZK_FMUL_NOINLINE: the multiplication becomes a real
// call there (smaller code, much faster to compile); the hot accumulate kernel inlines it.
#if defined(ZK_FMUL_NOINLINE) && defined(__CUDACC__)
#define ZK_FMUL_ATTR __host__ __device__ __noinline__
#else
#define ZK_FMUL_ATTR ZK_HD
#endif

template <class C> ZK_FMUL_ATTR void fmul(Mont<C>& r, const Mont<C>& a, const Mont<C>& b) {
  constexpr int n = C::N;
  static_assert(n % 2 == 0, "even limb count");
  uint32_t even[n], odd[n];
#pragma unroll
  for (int i = 0; i < n; i += 2) {
    detail::mont_step<C>(even, odd, a.v, b.v[i], i == 0);
    detail::mont_step<C>(odd, even, a.v, b.v[i + 1], false);
  }
  // last step used E = odd, O = even; E[0] == 0.  result = O + (E >> 32)  (< 2p)
  uint32_t t[n];
  t[0] = ptx::add_cc(even[0], odd[1]);
#pragma unroll
  for (int i = 1; i < n - 1; i++) t[i] = ptx::addc_cc(even[i], odd[i + 1]);
  t[n - 1] = ptx::addc(even[n - 1], 0);
  freduce_once(r, t);
}

namespace detail {

// t[0..2n) = a^2 as a plain integer.  Off-diagonal products a_i a_j (i < j) are summed once into two
// accumulators (64-bit digits at even resp. odd limb positions), doubled, and the squares a_i^2 are added:
// n(n+1)/2 = 78 limb products for n = 12 instead of 144.  Rows are taken in increasing i, so the carry out
// of a row's chain always lands in a limb no earlier row has used for a product (it holds 0, or a carry).
template <int n> ZK_HD void square_wide(uint32_t* t, const uint32_t* a) {
  uint32_t e[2 * n], o[2 * n];
#pragma unroll
  for (int k = 0; k < 2 * n; k++) { e[k] = 0; o[k] = 0; }
#pragma unroll
  for (int i = 0; i < n - 1; i++) {
    // positions i + j with j = i+1, i+3, ..: parity of (2i + 1) -> odd accumulator
    {
      int p = 2 * i + 1;
      o[p] = ptx::mad_lo_cc(a[i], a[i + 1], o[p]);
      o[p + 1] = ptx::madc_hi_cc(a[i], a[i + 1], o[p + 1]);
#pragma unroll
      for (int j = i + 3; j < n; j += 2) {
        p = i + j;
        o[p] = ptx::madc_lo_cc(a[i], a[j], o[p]);
        o[p + 1] = ptx::madc_hi_cc(a[i], a[j], o[p + 1]);
      }
      o[p + 2] = ptx::addc(o[p + 2], 0);
    }
    if (i + 2 < n) {  // j = i+2, i+4, ..: even accumulator
      int p = 2 * i + 2;
      e[p] = ptx::mad_lo_cc(a[i], a[i + 2], e[p]);
      e[p + 1] = ptx::madc_hi_cc(a[i], a[i + 2], e[p + 1]);
#pragma unroll
      for (int j = i + 4; j < n; j += 2) {
        p = i + j;
        e[p] = ptx::madc_lo_cc(a[i], a[j], e[p]);
        e[p + 1] = ptx::madc_hi_cc(a[i], a[j], e[p + 1]);
      }
      e[p + 2] = ptx::addc(e[p + 2], 0);
    }
  }
  // s = e + o (limb 0 and limb 2n-1 of the off-diagonal sum are zero), then t = 2 s
  uint32_t s[2 * n];
  s[0] = 0;
  s[1] = ptx::add_cc(e[1], o[1]);
#pragma unroll
  for (int k = 2; k < 2 * n - 1; k++) s[k] = ptx::addc_cc(e[k], o[k]);
  s[2 * n - 1] = ptx::addc(e[2 * n - 1], o[2 * n - 1]);
#pragma unroll
  for (int k = 2 * n - 1; k > 0; k--) s[k] = (s[k] << 1) | (s[k - 1] >> 31);
  // + diagonal: one carry chain over all 2n limbs
  t[0] = ptx::mad_lo_cc(a[0], a[0], s[0]);
  t[1] = ptx::madc_hi_cc(a[0], a[0], s[1]);
#pragma unroll
  for (int i = 1; i < n; i++) {
    t[2 * i] = ptx::madc_lo_cc(a[i], a[i], s[2 * i]);
    t[2 * i + 1] = ptx::madc_hi_cc(a[i], a[i], s[2 * i + 1]);
  }
}

// r = t / R mod p for a 2n-limb t < p * R (Montgomery reduction), result fully reduced.
// Same sliding even/odd window as fmul: each step cancels the low limb with m p and shifts; instead of a
// row a * b_i, the next limb of t enters at the top of the window.
template <class C> ZK_HD void mont_reduce_wide(Mont<C>& r, const uint32_t* t) {
  constexpr int n = C::N;
  uint32_t ev[n], od[n];
#pragma unroll
  for (int k = 0; k < n; k++) { ev[k] = t[k]; od[k] = 0; }
  uint32_t* E = ev;
  uint32_t* O = od;
#pragma unroll
  for (int i = 0; i < n; i++) {
    if (i > 0) {
      // shift: the old odd accumulator becomes even, the old even one (>> 64) becomes odd
      uint32_t* tmp = E; E = O; O = tmp;
      E[0] = ptx::add_cc(E[0], O[1]);
#pragma unroll
      for (int k = 0; k < n - 2; k++) O[k] = ptx::addc_cc(O[k + 2], 0);
      // incoming limb t[n + i - 1] sits at window limb n - 1 = odd digit (n-1, n): O[n-2], O[n-1]
      O[n - 2] = ptx::addc_cc(t[n + i - 1], 0);
      O[n - 1] = ptx::addc(0, 0);
    }
    uint32_t m = E[0] * C::INV;
    row_mad<n>(O, ModOdd<C>(), m);
    row_mad<n>(E, ModEven<C>(), m);
    O[n - 1] = ptx::addc(O[n - 1], 0);
  }
  // value = (E + O 2^32) / 2^32 + t[2n-1] 2^(32(n-1)), with E[0] == 0
  uint32_t u[n];
  u[0] = ptx::add_cc(O[0], E[1]);
#pragma unroll
  for (int k = 1; k < n - 1; k++) u[k] = ptx::addc_cc(O[k], E[k + 1]);
  u[n - 1] = ptx::addc(O[n - 1], 0);
  u[n - 1] += t[2 * n - 1];
  freduce_once(r, u);
}

}  // namespace detail

// Dedicated squaring: 78 + 144 limb products + 12 for the m_i, instead of 288 + 12 (22 % fewer multiplier slots).
template <class C> ZK_HD void fsqr(Mont<C>& r, const Mont<C>& a) {
#if (defined(ZK_FMUL_NOINLINE) && defined(__CUDACC__)) || defined(ZK_NO_FSQR)
  fmul(r, a, a);   // call-based translation units keep the single shared multiplication routine
#else
  uint32_t t[2 * C::N];
  detail::square_wide<C::N>(t, a.v);
  detail::mont_reduce_wide<C>(r, t);
#endif
}

// Two independent products advanced in lockstep (row by row), for the latency-bound kernels that run
// one warp per scheduler: the two carry chains overlap, so a pair costs little more than one product
// in wall time.  Always inlined.  r1/r2 may alias the inputs.
template <class C> ZK_HD void fmul2(Mont<C>& r1, const Mont<C>& a1, const Mont<C>& b1,
                                    Mont<C>& r2, const Mont<C>& a2, const Mont<C>& b2) {
  constexpr int n = C::N;
  uint32_t e1[n], o1[n], e2[n], o2[n];
#pragma unroll
  for (int i = 0; i < n; i += 2) {
    detail::mont_step<C>(e1, o1, a1.v, b1.v[i], i == 0);
    detail::mont_step<C>(e2, o2, a2.v, b2.v[i], i == 0);
    detail::mont_step<C>(o1, e1, a1.v, b1.v[i + 1], false);
    detail::mont_step<C>(o2, e2, a2.v, b2.v[i + 1], false);
  }
  uint32_t t1[n], t2[n];
  t1[0] = ptx::add_cc(e1[0], o1[1]);
#pragma unroll
  for (int i = 1; i < n - 1; i++) t1[i] = ptx::addc_cc(e1[i], o1[i + 1]);
  t1[n - 1] = ptx::addc(e1[n - 1], 0);
  t2[0] = ptx::add_cc(e2[0], o2[1]);
#pragma unroll
  for (int i = 1; i < n - 1; i++) t2[i] = ptx::addc_cc(e2[i], o2[i + 1]);
  t2[n - 1] = ptx::addc(e2[n - 1], 0);
  freduce_once(r1, t1);
  freduce_once(r2, t2);
}

// ---------------------------------------------------------------- conversions, inverse
// canonical limbs (< p) -> Montgomery form
template <class C> ZK_HD void fto_mont(Mont<C>& r, const uint32_t* canon) {
  Mont<C> a, r2;
#pragma unroll
  for (int i = 0; i < C::N; i++) { a.v[i] = canon[i]; r2.v[i] = C::r2(i); }
  fmul(r, a, r2);
}
// Montgomery form -> canonical limbs in [0, p)
template <class C> ZK_HD void ffrom_mont(uint32_t* canon, const Mont<C>& a) {
  Mont<C> one, r;
  fset_zero(one);
  one.v[0] = 1;
  fmul(r, a, one);
#pragma unroll
  for (int i = 0; i < C::N; i++) canon[i] = r.v[i];
}

// a^(p-2) by left-to-right square-and-multiply (Fermat).  Kept as the simple cross-check of finv.
template <class C> ZK_HD void finv_fermat(Mont<C>& r, const Mont<C>& a) {
  Mont<C> acc;
  fset_one(acc);
  for (int i = C::N - 1; i >= 0; i--) {
    uint32_t w = C::pm2(i);
    for (int b = 31; b >= 0; b--) {
      fsqr(acc, acc);
      if ((w >> b) & 1) fmul(acc, acc, a);
    }
  }
  r = acc;
}

namespace detail {
template <int n> ZK_HD bool limbs_is_one(const uint32_t* a) {
  uint32_t o = a[0] ^ 1u;
#pragma unroll
  for (int i = 1; i < n; i++) o |= a[i];
  return o == 0;
}
// a >>= 1 (a has n limbs plus the extra top bit `hi`)
template <int n> ZK_HD void limbs_shr1(uint32_t* a, uint32_t hi) {
#pragma unroll
  for (int i = 0; i < n - 1; i++) a[i] = (a[i] >> 1) | (a[i + 1] << 31);
  a[n - 1] = (a[n - 1] >> 1) | (hi << 31);
}
// x = x / 2 mod p  (x < p): x even -> x >> 1, else (x + p) >> 1
template <class C> ZK_HD void halve_mod(uint32_t* x) {
  constexpr int n = C::N;
  uint32_t odd = 0u - (x[0] & 1u);
  uint32_t t[n];
  t[0] = ptx::add_cc(x[0], C::p(0) & odd);
#pragma unroll
  for (int i = 1; i < n; i++) t[i] = ptx::addc_cc(x[i], C::p(i) & odd);
  uint32_t hi = ptx::addc(0, 0);
#pragma unroll
  for (int i = 0; i < n; i++) x[i] = t[i];
  limbs_shr1<n>(x, hi);
}
// a -= b (a >= b), plain integers
template <int n> ZK_HD void limbs_sub(uint32_t* a, const uint32_t* b) {
  a[0] = ptx::sub_cc(a[0], b[0]);
#pragma unroll
  for (int i = 1; i < n - 1; i++) a[i] = ptx::subc_cc(a[i], b[i]);
  a[n - 1] = ptx::subc(a[n - 1], b[n - 1]);
}
template <int n> ZK_HD bool limbs_ge(const uint32_t* a, const uint32_t* b) {
  uint32_t t = ptx::sub_cc(a[0], b[0]);
#pragma unroll
  for (int i = 1; i < n; i++) t = ptx::subc_cc(a[i], b[i]);
  (void)t;
  return ptx::subc(0, 0) == 0;  // no borrow
}
}  // namespace detail

// Modular inverse by the binary extended Euclid (shift / subtract only, ~760 short carry-chain steps of ~250
// cycles).  Kept as the second cross-check of finv (tests/test_host_emu_field.py); Montgomery in / out, inv(0) = 0.
template <class C> ZK_HD void finv_euclid(Mont<C>& r, const Mont<C>& a) {
  constexpr int n = C::N;
  if (fis_zero(a)) { fset_zero(r); return; }
  uint32_t u[n], v[n];
  Mont<C> x1, x2;
#pragma unroll
  for (int i = 0; i < n; i++) { u[i] = a.v[i]; v[i] = C::p(i); x1.v[i] = 0; x2.v[i] = 0; }
  x1.v[0] = 1;
  while (!detail::limbs_is_one<n>(u) && !detail::limbs_is_one<n>(v)) {
    while (!(u[0] & 1)) { detail::limbs_shr1<n>(u, 0); detail::halve_mod<C>(x1.v); }
    while (!(v[0] & 1)) { detail::limbs_shr1<n>(v, 0); detail::halve_mod<C>(x2.v); }
    if (detail::limbs_ge<n>(u, v)) { detail::limbs_sub<n>(u, v); fsub(x1, x1, x2); }
    else { detail::limbs_sub<n>(v, u); fsub(x2, x2, x1); }
  }
  // plain inverse of the Montgomery representative aR is a^-1 R^-1; times R^3 (Montgomery) = a^-1 R
  Mont<C> r3, t = detail::limbs_is_one<n>(u) ? x1 : x2;
#pragma unroll
  for (int i = 0; i < n; i++) r3.v[i] = C::r3(i);
  fmul(r, t, r3);
}

// ---- finv: division steps in batches of 30 (Bernstein-Yang "safegcd", variable-time form) ------------------
// Exactly ONE thread inverts at the end of an MSM, so what counts is the length of its dependent instruction
// chain.  The binary Euclid above touches all limbs of four numbers at every one of its ~760 steps; here 30
// steps run on the low 32 bits of (f, g) alone and leave a 2x2 integer matrix t with 2^30 (f, g)' = t (f, g),
// which is then applied once to the full numbers (f, g) and, modulo p, to the cofactors (d, e):
// ~27 batches x (13-limb multiply-accumulate passes) ~ 20 us instead of ~100 us.  Numbers are held in 30-bit
// signed limbs so the division by 2^30 is a limb drop and every product is one 32x32->64 signed multiply-add.
// Invariants: d x = f, e x = g (mod p); |f|, |g| <= p; d, e in (-2p, p).  Ends when g = 0: f = +-1, d = +-x^-1.
namespace detail {
static constexpr uint32_t M30 = 0x3fffffffu;

ZK_HD int ctz32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  return __ffs((int)x) - 1;
#else
  return __builtin_ctz(x);
#endif
}

template <int NL, int M> ZK_HD void to_s30(const uint32_t* a, int32_t* out) {
#pragma unroll
  for (int i = 0; i < M; i++) {
    const int bit = 30 * i, w = bit / 32, s = bit % 32;
    uint32_t x = w < NL ? a[w] >> s : 0u;
    if (s > 2 && w + 1 < NL) x |= a[w + 1] << (32 - s);
    out[i] = (int32_t)(x & M30);
  }
}
// limbs already in [0, 2^30)
template <int NL, int M> ZK_HD void from_s30(const int32_t* v, uint32_t* a) {
#pragma unroll
  for (int w = 0; w < NL; w++) {
    const int bit = 32 * w, i = bit / 30, s = bit % 30;   // s is even and <= 28
    uint32_t x = (uint32_t)v[i] >> s;
    if (i + 1 < M) x |= (uint32_t)v[i + 1] << (30 - s);
    a[w] = x;
  }
}

// 30 division steps on the low words; t = (u, v, q, r) with 2^30 f' = u f + v g, 2^30 g' = q f + r g.
// eta = -delta of the paper.  g is made even by adding a multiple of f (both odd), after swapping (f, g) <- (g, -f)
// when eta < 0.
ZK_HD int32_t divsteps30(int32_t eta, uint32_t f, uint32_t g, int32_t* t) {
  uint32_t u = 1, v = 0, q = 0, r = 1;
  int i = 30;
  for (;;) {
    int zeros = ctz32(g | (0xffffffffu << i));
    g >>= zeros; u <<= zeros; v <<= zeros;
    eta -= zeros; i -= zeros;
    if (i == 0) break;
    if (eta < 0) {
      uint32_t x;
      eta = -eta;
      x = f; f = g; g = 0u - x;
      x = u; u = q; q = 0u - x;
      x = v; v = r; r = 0u - x;
    }
    // cancel up to 6 low bits of g at once: w = -g / f mod 2^limit (f (f^2 - 2) = -1/f mod 64), which is `limit`
    // consecutive steps of the "add f" kind -- allowed while eta stays >= 0, i.e. limit <= eta + 1
    int limit = eta + 1 > i ? i : eta + 1;
    uint32_t w = (f * g * (f * f - 2u)) & (0xffffffffu >> (32 - limit)) & 63u;
    g += f * w; q += u * w; r += v * w;
  }
  t[0] = (int32_t)u; t[1] = (int32_t)v; t[2] = (int32_t)q; t[3] = (int32_t)r;
  return eta;
}

// (f, g) <- t (f, g) / 2^30   (exact)
template <int M> ZK_HD void update_fg30(int32_t* f, int32_t* g, const int32_t* t) {
  const int64_t u = t[0], v = t[1], q = t[2], r = t[3];
  int64_t cf = u * f[0] + v * g[0], cg = q * f[0] + r * g[0];
  cf >>= 30; cg >>= 30;
#pragma unroll
  for (int i = 1; i < M; i++) {
    cf += u * f[i] + v * g[i];
    cg += q * f[i] + r * g[i];
    f[i - 1] = (int32_t)((uint32_t)cf & M30); cf >>= 30;
    g[i - 1] = (int32_t)((uint32_t)cg & M30); cg >>= 30;
  }
  f[M - 1] = (int32_t)cf;
  g[M - 1] = (int32_t)cg;
}

// (d, e) <- t (d, e) / 2^30 mod p: multiples md p, me p are added so that the low 30 bits cancel
// (minv = p^-1 mod 2^30); keeps d, e in (-2p, p).
template <int M> ZK_HD void update_de30(int32_t* d, int32_t* e, const int32_t* t, const int32_t* m, uint32_t minv) {
  const int64_t u = t[0], v = t[1], q = t[2], r = t[3];
  const int32_t sd = d[M - 1] >> 31, se = e[M - 1] >> 31;
  int32_t md = (t[0] & sd) + (t[1] & se), me = (t[2] & sd) + (t[3] & se);
  int64_t cd = u * d[0] + v * e[0], ce = q * d[0] + r * e[0];
  md -= (int32_t)((minv * (uint32_t)cd + (uint32_t)md) & M30);
  me -= (int32_t)((minv * (uint32_t)ce + (uint32_t)me) & M30);
  cd += (int64_t)m[0] * md;
  ce += (int64_t)m[0] * me;
  cd >>= 30; ce >>= 30;
#pragma unroll
  for (int i = 1; i < M; i++) {
    cd += u * d[i] + v * e[i] + (int64_t)m[i] * md;
    ce += q * d[i] + r * e[i] + (int64_t)m[i] * me;
    d[i - 1] = (int32_t)((uint32_t)cd & M30); cd >>= 30;
    e[i - 1] = (int32_t)((uint32_t)ce & M30); ce >>= 30;
  }
  d[M - 1] = (int32_t)cd;
  e[M - 1] = (int32_t)ce;
}

template <int M> ZK_HD void carry30(int32_t* x) {
#pragma unroll
  for (int i = 0; i < M - 1; i++) { x[i + 1] += x[i] >> 30; x[i] &= (int32_t)M30; }
}
// x in (-2p, p)  ->  sign * x mod p in [0, p), limbs in [0, 2^30)
template <int M> ZK_HD void normalize30(int32_t* x, int32_t sign, const int32_t* m) {
  int32_t add = x[M - 1] >> 31;
  const int32_t neg = sign >> 31;
#pragma unroll
  for (int i = 0; i < M; i++) { x[i] += m[i] & add; x[i] = (x[i] ^ neg) - neg; }
  carry30<M>(x);
  add = x[M - 1] >> 31;
#pragma unroll
  for (int i = 0; i < M; i++) x[i] += m[i] & add;
  carry30<M>(x);
}
}  // namespace detail

// Replaces the extended-Euclid inverse of prime_field_elem.rs:379-432; same value since p is prime.
// Montgomery in, Montgomery out; inv(0) = 0.
template <class C> ZK_HD void finv(Mont<C>& r, const Mont<C>& a) {
  constexpr int NL = C::N, M = NL * 32 / 30 + 1;
  int32_t m[M], f[M], g[M], d[M], e[M];
  uint32_t pl[NL];
#pragma unroll
  for (int i = 0; i < NL; i++) pl[i] = C::p(i);
  detail::to_s30<NL, M>(pl, m);
  detail::to_s30<NL, M>(a.v, g);
#pragma unroll
  for (int i = 0; i < M; i++) { f[i] = m[i]; d[i] = 0; e[i] = 0; }
  e[0] = 1;
  const uint32_t minv = (0u - C::INV) & detail::M30;
  int32_t eta = -1;
  for (int batch = 0; batch < 64; batch++) {   // <= (49 bits + 57) / 17 steps: 37 batches for 381 bits
    int32_t nz = 0;
#pragma unroll
    for (int i = 0; i < M; i++) nz |= g[i];
    if (nz == 0) break;
    int32_t t[4];
    eta = detail::divsteps30(eta, (uint32_t)f[0] | ((uint32_t)f[1] << 30), (uint32_t)g[0] | ((uint32_t)g[1] << 30), t);
    detail::update_de30<M>(d, e, t, m, minv);
    detail::update_fg30<M>(f, g, t);
  }
  detail::normalize30<M>(d, f[M - 1], m);
  // plain inverse of the Montgomery representative aR is a^-1 R^-1; times R^3 (Montgomery) = a^-1 R
  Mont<C> r3, x;
  detail::from_s30<NL, M>(d, x.v);
#pragma unroll
  for (int i = 0; i < NL; i++) r3.v[i] = C::r3(i);
  fmul(r, x, r3);
}

}  // namespace zk
