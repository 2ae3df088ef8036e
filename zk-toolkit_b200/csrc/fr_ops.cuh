// Fr vector kernels next to the MSM: the witness aggregation of the Groth16 prover.
//
// Reference: `for i in 0..=m { ... ui[i] ... * ai }` (zk/w_trusted_setup/groth16/zktoolkit_based/prover.rs:108-117)
// and QAP::build_p's `v += &self.vi[i] * wit` (qap/qap.rs:99-109, scalar * polynomial = polynomial.rs:351-361):
// out_j = sum_i a_i * poly_i[j] mod r.  The per-wire polynomials are dense (Lagrange interpolants over 1..n,
// qap.rs:33-97), so the input is a dense n_wires x n matrix of canonical Fr limbs.
#pragma once
#include "fp.cuh"

namespace zk {

#if defined(__CUDA_ARCH__)
ZK_D void zk_atomic_or_u32(uint32_t* p, uint32_t v) { atomicOr(p, v); }
#else
inline void zk_atomic_or_u32(uint32_t* p, uint32_t v) { *p |= v; }
#endif

// wires -> Montgomery form once, so that mont(a_i R, c) = a_i c comes out canonical
struct FrToMont {
  static const char* name() { return "fr_to_mont"; }
  static ZK_HD void run(uint32_t tid, uint32_t n, const uint32_t* in, Fr* out) {
    if (tid >= n) return;
    uint32_t a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = in[(size_t)tid * 8 + i];
    Fr r;
    fto_mont(r, a);
    out[tid] = r;
  }
};

// thread j: out[j] = sum_i wires[i] * polys[i][j]; *err |= 1 if any coefficient is not reduced (>= r)
struct FrAggregate {
  static const char* name() { return "fr_aggregate"; }
  static ZK_HD void run(uint32_t tid, uint32_t n_wires, uint32_t n, const Fr* wires_mont, const uint32_t* polys,
                        uint32_t* out) {
    if (tid >= n) return;
    Fr acc;
    fset_zero(acc);
    for (uint32_t i = 0; i < n_wires; i++) {
      Fr c, a = wires_mont[i], t;
      const uint32_t* src = polys + ((size_t)i * n + tid) * 8;
#pragma unroll
      for (int k = 0; k < 8; k++) c.v[k] = src[k];
      fmul(t, a, c);          // (a R) c / R = a c, fully reduced
      fadd(acc, acc, t);
    }
#pragma unroll
    for (int k = 0; k < 8; k++) out[(size_t)tid * 8 + k] = acc.v[k];
  }
};

// ---------------------------------------------------------------- scalar vectors of one Groth16 proof
// Prover::prove (groth16/zktoolkit_based/prover.rs:96-147) as three MSMs (zkmsm_groth16_prove):
//   A = alpha + sum_j u_j [x^j]_1 + r delta                        scalars  u ++ [1, r]
//   B = beta_2 + sum_j v_j [x^j]_2 + s delta_2                     scalars  v ++ [1, s]
//   C = sum a_i uvw_wit_i + sum h_j [x^j t/delta]_1 + s A + r B_g1 - r s delta       (:128-145)
//     = sum a_i uvw_wit_i + sum h_j [x^j t/delta]_1 + sum_j (s u_j + r v_j) [x^j]_1 + s alpha + r beta + (r s) delta
// (expand A and B_g1 = beta + sum v_j [x^j]_1 + s delta): one MSM with scalars wit ++ h ++ (s u + r v) ++ [s, r, r s],
// so B_g1 and the two 255-bit scalar multiplications s A, r B_g1 of the reference are never formed.
// This body fills what is not a plain copy of the caller's vectors: thread j < n writes (s u_j + r v_j) mod r,
// thread n the trailing slots.  rs = r | s (canonical, < r, else *err |= 1).
struct Groth16Scalars {
  static const char* name() { return "groth16_scalars"; }
  static ZK_HD bool below_r(const uint32_t* a) {
    uint32_t t = ptx::sub_cc(a[0], FR_P[0]);
#pragma unroll
    for (int i = 1; i < 8; i++) t = ptx::subc_cc(a[i], FR_P[i]);
    (void)t;
    return ptx::subc(0, 0) != 0;   // borrow: a < r
  }
  // write_b_tail = 0: sv's trailing [1, s] was already written by Groth16TailB (the G2 MSM started on it)
  static ZK_HD void run(uint32_t tid, uint32_t n, const uint32_t* rs, uint32_t* su, uint32_t* sv, uint32_t* sc3, uint32_t write_b_tail,
                        uint32_t* err) {
    if (tid > n) return;
    uint32_t rc[8], sc[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { rc[i] = rs[i]; sc[i] = rs[8 + i]; }
    Fr rm, sm;
    fto_mont(rm, rc);
    fto_mont(sm, sc);
    if (tid < n) {
      Fr u, v, a, b;
#pragma unroll
      for (int i = 0; i < 8; i++) { u.v[i] = su[(size_t)tid * 8 + i]; v.v[i] = sv[(size_t)tid * 8 + i]; }
      fmul(a, sm, u);   // (s R) u / R = s u, fully reduced
      fmul(b, rm, v);
      fadd(a, a, b);
#pragma unroll
      for (int i = 0; i < 8; i++) sc3[(size_t)tid * 8 + i] = a.v[i];
      return;
    }
    if (!below_r(rc) || !below_r(sc)) zk_atomic_or_u32(err, 1u);
    Fr prod, sone;
    fset_zero(sone);
    sone.v[0] = 1;
    fmul(prod, rm, sm);          // r s R
    fmul(prod, prod, sone);      // r s (canonical)
    uint32_t* a_tail = su + (size_t)n * 8;
    uint32_t* b_tail = sv + (size_t)n * 8;
    uint32_t* c_tail = sc3 + (size_t)n * 8;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      a_tail[i] = i == 0 ? 1u : 0u; a_tail[8 + i] = rc[i];
      if (write_b_tail) { b_tail[i] = i == 0 ? 1u : 0u; b_tail[8 + i] = sc[i]; }
      c_tail[i] = sc[i]; c_tail[8 + i] = rc[i]; c_tail[16 + i] = prod.v[i];
    }
  }
};

// sv's trailing slots alone, so that B's MSM can start as soon as v has arrived (before u, h and the witness)
struct Groth16TailB {
  static const char* name() { return "groth16_tail_b"; }
  static ZK_HD void run(uint32_t tid, uint32_t n, const uint32_t* rs, uint32_t* sv) {
    if (tid != 0) return;
    uint32_t* b_tail = sv + (size_t)n * 8;
#pragma unroll
    for (int i = 0; i < 8; i++) { b_tail[i] = i == 0 ? 1u : 0u; b_tail[8 + i] = rs[8 + i]; }
  }
};

// ---------------------------------------------------------------- quotient polynomial h = (u v - w) / t
// Reference: Prover::new (groth16/zktoolkit_based/prover.rs:64-71): p = qap.build_p(witness) = u*v - w
// (qap.rs:99-112, Polynomial::multiply_by polynomial.rs:173-190), t = prod_{k=1..n} (x - k) (QAP::build_t,
// qap.rs:115-135), h = p.divide_by(t) (polynomial long division, polynomial.rs:204-238; "p should be divisible
// by t").  Same schoolbook algorithms, parallel over coefficients; all values Montgomery Fr on the device.

struct FrVecToMont {   // in: n x 8 canonical words -> out: Fr (Montgomery); entries beyond n_in are zero
  static const char* name() { return "fr_vec_to_mont"; }
  static ZK_HD void run(uint32_t tid, uint32_t n_out, uint32_t n_in, const uint32_t* in, Fr* out) {
    if (tid >= n_out) return;
    Fr r;
    if (tid < n_in) {
      uint32_t a[8];
#pragma unroll
      for (int i = 0; i < 8; i++) a[i] = in[(size_t)tid * 8 + i];
      fto_mont(r, a);
    } else fset_zero(r);
    out[tid] = r;
  }
};

struct FrPolyMulSub {   // p[k] = sum_{i+j=k} u[i] v[j] - w[k],  k < 2n-1  (u, v, w have n coefficients)
  static const char* name() { return "fr_poly_mul_sub"; }
  static ZK_HD void run(uint32_t tid, uint32_t n, const Fr* u, const Fr* v, const Fr* w, Fr* p) {
    if (tid >= 2 * n - 1) return;
    uint32_t lo = tid >= n ? tid - n + 1 : 0, hi = tid < n ? tid : n - 1;
    Fr acc;
    fset_zero(acc);
    for (uint32_t i = lo; i <= hi; i++) {
      Fr a = u[i], b = v[tid - i], t;
      fmul(t, a, b);
      fadd(acc, acc, t);
    }
    if (tid < n) { Fr c = w[tid]; fsub(acc, acc, c); }
    p[tid] = acc;
  }
};

// t_k(x) = t_{k-1}(x) (x - k): thread j <= k writes coefficient j; t_0 = 1.  k_mont = k in Montgomery form.
struct FrTStep {
  static const char* name() { return "fr_t_step"; }
  static ZK_HD void run(uint32_t tid, uint32_t k, Fr k_mont, const Fr* t_old, Fr* t_new) {
    if (tid > k) return;
    Fr lower, cur, prod, r;
    if (tid > 0) lower = t_old[tid - 1]; else fset_zero(lower);
    if (tid < k) cur = t_old[tid]; else fset_zero(cur);
    fmul(prod, cur, k_mont);
    fsub(r, lower, prod);
    t_new[tid] = r;
  }
};

// one step of the long division by the monic t (degree n): q = p[d + n]; h[d] = q; p[d + j] -= q t[j], j < n
struct FrDivStep {
  static const char* name() { return "fr_div_step"; }
  static ZK_HD void run(uint32_t tid, uint32_t n, uint32_t d, const Fr* t, Fr* p, Fr* h) {
    if (tid >= n) return;
    Fr q = p[d + n], tj = t[tid], prod, cur = p[d + tid];
    fmul(prod, q, tj);
    fsub(cur, cur, prod);
    p[d + tid] = cur;
    if (tid == 0) h[d] = q;
  }
};

// h (Montgomery) -> canonical words; *nonzero_rem |= 1 if any of the low n coefficients of p is not zero
struct FrQuotientOut {
  static const char* name() { return "fr_quotient_out"; }
  static ZK_HD void run(uint32_t tid, uint32_t n, const Fr* h, const Fr* p, uint32_t* out, uint32_t* nonzero_rem) {
    if (tid >= n) return;
    if (!fis_zero(p[tid])) zk_atomic_or_u32(nonzero_rem, 1u);
    if (tid < n - 1) ffrom_mont(out + (size_t)tid * 8, h[tid]);
  }
};

}  // namespace zk
