// Fr vector kernels next to the MSM: the witness aggregation of the Groth16 prover.
//
// Reference: `for i in 0..=m { ... ui[i] ... * ai }` (zk/w_trusted_setup/groth16/zktoolkit_based/prover.rs:108-117)
// and QAP::build_p's `v += &self.vi[i] * wit` (qap/qap.rs:99-109, scalar * polynomial = polynomial.rs:351-361):
// out_j = sum_i a_i * poly_i[j] mod r.  The per-wire polynomials are dense (Lagrange interpolants over 1..n,
// qap.rs:33-97), so the input is a dense n_wires x n matrix of canonical Fr limbs.
#pragma once
#include "fp.cuh"

namespace zk {

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

}  // namespace zk
