// Quotient polynomial h = (u v - w) / t in O(n log n): number-theoretic transforms over Fr.
//
// Reference: Prover::new (groth16/zktoolkit_based/prover.rs:64-71): p = qap.build_p(witness) = u*v - w (qap.rs:99-112),
// t = prod_{k=1..n} (x - k) (QAP::build_t, qap.rs:115-135), h = p.divide_by(t) (polynomial.rs:204-238), panic
// "p should be divisible by t" otherwise.  The reference multiplies and divides schoolbook-style (O(n^2), kept in
// fr_ops.cuh for small n and as the cross-check); the values are exact field elements, so any algorithm gives the
// same coefficients.  Here:
//   * r - 1 = 2^32 * odd, so Fr has transforms of every power-of-two size up to 2^32 (FR_ROOT_2_32).
//   * t is built once per n by a product tree over the linear factors (monic polynomials, leading 1 implicit):
//     (x^d + a)(x^d + b) = x^2d + (a + b) x^d + a b, the product a b by a batched transform of size 2d.
//   * exact division by reversal: rev(h) = rev(p) * rev(t)^-1 mod x^(n-1); the power-series inverse of rev(t) comes
//     from Newton's iteration I <- I (2 - rev(t) I), also once per n.
//   * per proof: 7 transforms of size M = 2^ceil(log2 2n): u, v, p^-1, rev(p), rev(h)^-1, h, (h t)^-1; the last two
//     verify p == h t coefficient by coefficient (the reference's divisibility panic becomes *exact = 0).
// Transforms are radix-2, forward = decimation in frequency (natural in, bit-reversed out), inverse = decimation in
// time (bit-reversed in, natural out): products are taken in bit-reversed order and no permutation pass exists.
// Every body is a plain function of (tid, args) like the MSM stages, so tests/host_emu runs the same pipeline on the CPU.
#pragma once
#include "fr_ops.cuh"

namespace zk {

// consts[0] = w (primitive M-th root), consts[1] = w^-1, consts[2] = M^-1, all Montgomery; one thread
struct FrNttSetup {
  static const char* name() { return "fr_ntt_setup"; }
  static ZK_HD void run(uint32_t tid, uint32_t logm, Fr* consts) {
    if (tid) return;
    Fr w, m;
#pragma unroll
    for (int i = 0; i < 8; i++) w.v[i] = FR_ROOT_2_32[i];
    for (uint32_t i = logm; i < 32; i++) fmul(w, w, w);
    consts[0] = w;
    finv(consts[1], w);
    fset_one(m);
    for (uint32_t i = 0; i < logm; i++) fadd(m, m, m);
    finv(consts[2], m);
  }
};

// tw[i] = base^i, i < count (square and multiply over the bits of i)
struct FrPowTable {
  static const char* name() { return "fr_pow_table"; }
  static ZK_HD void run(uint32_t tid, uint32_t count, const Fr* base, Fr* tw) {
    if (tid >= count) return;
    Fr r, b = *base;
    fset_one(r);
    for (uint32_t e = tid; e; e >>= 1) {
      if (e & 1u) fmul(r, r, b);
      fmul(b, b, b);
    }
    tw[tid] = r;
  }
};

// One butterfly stage over an array that is a concatenation of independent transforms; blocks of `len` elements,
// thread t = butterfly (block t / half, position t % half).  tw[k] = w_M^k (forward) or w_M^-k (inverse), k < M / 2;
// the root of a block of length len is w_M^(M / len).
struct FrNttDif {   // forward stage: (x, y) -> (x + y, (x - y) w^pos)
  static const char* name() { return "fr_ntt_dif"; }
  static ZK_HD void run(uint32_t tid, uint32_t total_half, uint32_t len, uint32_t m_over_len, const Fr* tw, Fr* a) {
    if (tid >= total_half) return;
    const uint32_t half = len / 2, pos = tid % half;
    const size_t i = (size_t)(tid / half) * len + pos;
    Fr x = a[i], y = a[i + half], s, d;
    fadd(s, x, y);
    fsub(d, x, y);
    if (pos) { Fr w = tw[(size_t)pos * m_over_len]; fmul(d, d, w); }
    a[i] = s;
    a[i + half] = d;
  }
};
struct FrNttDit {   // inverse stage: (x, y) -> (x + y w^-pos, x - y w^-pos)
  static const char* name() { return "fr_ntt_dit"; }
  static ZK_HD void run(uint32_t tid, uint32_t total_half, uint32_t len, uint32_t m_over_len, const Fr* twi, Fr* a) {
    if (tid >= total_half) return;
    const uint32_t half = len / 2, pos = tid % half;
    const size_t i = (size_t)(tid / half) * len + pos;
    Fr x = a[i], y = a[i + half], s, d;
    if (pos) { Fr w = twi[(size_t)pos * m_over_len]; fmul(y, y, w); }
    fadd(s, x, y);
    fsub(d, x, y);
    a[i] = s;
    a[i + half] = d;
  }
};
// a[i] *= M^-1 * 2^shift  (= m^-1 for transforms of size m = M / 2^shift)
struct FrScalePow2 {
  static const char* name() { return "fr_scale"; }
  static ZK_HD void run(uint32_t tid, uint32_t total, uint32_t shift, const Fr* consts, Fr* a) {
    if (tid >= total) return;
    Fr s = consts[2], x = a[tid];
    for (uint32_t i = 0; i < shift; i++) fadd(s, s, s);
    fmul(x, x, s);
    a[tid] = x;
  }
};
struct FrPointMul {   // out[i] = a[i] b[i]
  static const char* name() { return "fr_point_mul"; }
  static ZK_HD void run(uint32_t tid, uint32_t total, const Fr* a, const Fr* b, Fr* out) {
    if (tid >= total) return;
    Fr x = a[tid], y = b[tid], r;
    fmul(r, x, y);
    out[tid] = r;
  }
};
// dst[i] = i < k ? src[i] : 0, i < total
struct FrCopyPad {
  static const char* name() { return "fr_copy_pad"; }
  static ZK_HD void run(uint32_t tid, uint32_t total, uint32_t k, const Fr* src, Fr* dst) {
    if (tid >= total) return;
    Fr x;
    if (tid < k) x = src[tid]; else fset_zero(x);
    dst[tid] = x;
  }
};

// ---- product tree for T(x) = x^pad * prod_{k=1..n} (x - k), pad = N2 - n: monic of degree N2, low N2 coefficients
struct FrTreeLeaves {   // leaf i = x - (i + 1) for i < n, x for the padding: low part -(i+1) or 0
  static const char* name() { return "fr_tree_leaves"; }
  static ZK_HD void run(uint32_t tid, uint32_t n2, uint32_t n, Fr* poly) {
    if (tid >= n2) return;
    Fr r;
    if (tid < n) {
      uint32_t k[8] = {tid + 1, 0, 0, 0, 0, 0, 0, 0};
      fto_mont(r, k);
      fneg(r, r);
    } else fset_zero(r);
    poly[tid] = r;
  }
};
// pair q of nodes of degree d: X[q 4d .. +2d) = a padded, X[q 4d + 2d .. +2d) = b padded
struct FrTreeExpand {
  static const char* name() { return "fr_tree_expand"; }
  static ZK_HD void run(uint32_t tid, uint32_t total, uint32_t d, const Fr* poly, Fr* x) {
    if (tid >= total) return;                      // total = 2 * N2
    const uint32_t q = tid / (4 * d), r = tid % (4 * d), which = r / (2 * d), i = r % (2 * d);
    Fr v;
    if (i < d) v = poly[(size_t)q * 2 * d + which * d + i]; else fset_zero(v);
    x[tid] = v;
  }
};
struct FrTreeMul {   // Y[q 2d + i] = X[q 4d + i] * X[q 4d + 2d + i]
  static const char* name() { return "fr_tree_mul"; }
  static ZK_HD void run(uint32_t tid, uint32_t total, uint32_t d, const Fr* x, Fr* y) {
    if (tid >= total) return;                      // total = N2
    const uint32_t q = tid / (2 * d), i = tid % (2 * d);
    Fr a = x[(size_t)q * 4 * d + i], b = x[(size_t)q * 4 * d + 2 * d + i], r;
    fmul(r, a, b);
    y[tid] = r;
  }
};
struct FrTreeCombine {   // out[q 2d + i] = y[q 2d + i] + (i >= d ? a[i - d] + b[i - d] : 0)
  static const char* name() { return "fr_tree_combine"; }
  static ZK_HD void run(uint32_t tid, uint32_t total, uint32_t d, const Fr* y, const Fr* poly, Fr* out) {
    if (tid >= total) return;
    const uint32_t q = tid / (2 * d), i = tid % (2 * d);
    Fr r = y[tid];
    if (i >= d) {
      Fr a = poly[(size_t)q * 2 * d + (i - d)], b = poly[(size_t)q * 2 * d + d + (i - d)];
      fadd(r, r, a);
      fadd(r, r, b);
    }
    out[tid] = r;
  }
};
// t_j = T_(j + pad), j <= n (the leading 1 explicit); t has n + 1 coefficients, zero padded to total
struct FrTreeToT {
  static const char* name() { return "fr_tree_to_t"; }
  static ZK_HD void run(uint32_t tid, uint32_t total, uint32_t n2, uint32_t n, const Fr* tree, Fr* t) {
    if (tid >= total) return;
    Fr r;
    if (tid < n) r = tree[tid + (n2 - n)];
    else if (tid == n) fset_one(r);
    else fset_zero(r);
    t[tid] = r;
  }
};
// f = rev_n(t) mod x^k: f_i = t_(n - i) for i < min(k, n + 1), zero padded to total
struct FrRevT {
  static const char* name() { return "fr_rev_t"; }
  static ZK_HD void run(uint32_t tid, uint32_t total, uint32_t k, uint32_t n, const Fr* t, Fr* f) {
    if (tid >= total) return;
    Fr r;
    if (tid < k && tid <= n) r = t[n - tid]; else fset_zero(r);
    f[tid] = r;
  }
};
// Newton: g = 2 - e mod x^k (e = f I mod x^k), zero padded to total
struct FrTwoMinus {
  static const char* name() { return "fr_two_minus"; }
  static ZK_HD void run(uint32_t tid, uint32_t total, uint32_t k, const Fr* e, Fr* g) {
    if (tid >= total) return;
    Fr r;
    if (tid < k) {
      r = e[tid];
      fneg(r, r);
      if (tid == 0) { Fr one; fset_one(one); fadd(r, r, one); fadd(r, r, one); }
    } else fset_zero(r);
    g[tid] = r;
  }
};

// ---- per proof
struct FrSubW {   // p[i] -= w[i], i < n
  static const char* name() { return "fr_sub_w"; }
  static ZK_HD void run(uint32_t tid, uint32_t n, const Fr* w, Fr* p) {
    if (tid >= n) return;
    Fr a = p[tid], b = w[tid];
    fsub(a, a, b);
    p[tid] = a;
  }
};
struct FrRevTop {   // a_i = p_(2n - 2 - i), i < n - 1; zero padded to total
  static const char* name() { return "fr_rev_top"; }
  static ZK_HD void run(uint32_t tid, uint32_t total, uint32_t n, const Fr* p, Fr* a) {
    if (tid >= total) return;
    Fr r;
    if (tid + 1 < n) r = p[2 * n - 2 - tid]; else fset_zero(r);
    a[tid] = r;
  }
};
struct FrExtractH {   // h_j = q_(n - 2 - j), j < n - 1: canonical words out, Montgomery copy zero padded to total
  static const char* name() { return "fr_extract_h"; }
  static ZK_HD void run(uint32_t tid, uint32_t total, uint32_t n, const Fr* q, Fr* h, uint32_t* out) {
    if (tid >= total) return;
    Fr r;
    if (tid + 1 < n) {
      r = q[n - 2 - tid];
      ffrom_mont(out + (size_t)tid * 8, r);
    } else fset_zero(r);
    h[tid] = r;
  }
};
struct FrCheckEqual {   // *flag |= 1 unless a[i] == b[i], i < total
  static const char* name() { return "fr_check_equal"; }
  static ZK_HD void run(uint32_t tid, uint32_t total, const Fr* a, const Fr* b, uint32_t* flag) {
    if (tid >= total) return;
    if (!feq(a[tid], b[tid])) zk_atomic_or_u32(flag, 1u);
  }
};

// ------------------------------------------------------------------------------------------------ host pipeline
struct FrNttPlan {
  uint32_t n, logm, m;   // m = 2^logm >= 2n, m >= 4
  ZK_HD static FrNttPlan make(uint32_t n) {
    FrNttPlan p;
    p.n = n;
    p.logm = 2;
    while ((1ull << p.logm) < 2ull * n) p.logm++;
    p.m = 1u << p.logm;
    return p;
  }
};
// cached per n (owned by the caller): tw, twi (m/2 each), consts (4), that, ihat (m each), t (m)
struct FrNttTables { Fr *tw, *twi, *consts, *that, *ihat, *t; };
static inline size_t fr_ntt_table_elems(const FrNttPlan& p) { return (size_t)p.m / 2 * 2 + 4 + 3 * (size_t)p.m; }
static inline void fr_ntt_tables_at(FrNttTables& tb, const FrNttPlan& p, Fr* base) {
  tb.tw = base; tb.twi = tb.tw + p.m / 2; tb.consts = tb.twi + p.m / 2; tb.that = tb.consts + 4; tb.ihat = tb.that + p.m;
  tb.t = tb.ihat + p.m;
}

// A transform of 2^q points is cut into passes of up to 8 stages: block lengths 2^q .. 2^9 in strided passes, the
// last 8 stages (contiguous blocks of 256) in one local pass.  Exec::ntt_fused runs a pass through shared memory
// (tu_fr_ntt.cu) or reports that it cannot (CPU emulation, small or oddly shaped batches): then the stages of
// the pass are launched one by one.
struct FrNttPass { uint32_t len, k; };
static inline int fr_ntt_passes(uint32_t size, FrNttPass* out) {
  int n = 0;
  uint32_t len = size;
  while (len >= 2) {
    uint32_t lg = 0;
    while ((1u << lg) < len) lg++;
    uint32_t k = lg <= 8 ? lg : (lg - 8 < 8 ? lg - 8 : 8);
    out[n].len = len; out[n].k = k; n++;
    len >>= k;
  }
  return n;
}
// forward transform of `total / size` concatenated blocks of `size` (natural -> bit-reversed)
template <class Exec> void fr_ntt_forward(Exec& ex, const FrNttPlan& p, const FrNttTables& tb, Fr* a, uint32_t total, uint32_t size) {
  FrNttPass pass[8];
  const int np = fr_ntt_passes(size, pass);
  for (int i = 0; i < np; i++) {
    if (ex.ntt_fused(false, a, total, pass[i].len, pass[i].k, (const Fr*)tb.tw, p.m)) continue;
    for (uint32_t s = 0; s < pass[i].k; s++) {
      uint32_t len = pass[i].len >> s;
      ex.template launch<FrNttDif>(total / 2, total / 2, len, p.m / len, (const Fr*)tb.tw, a);
    }
  }
}
// inverse (bit-reversed -> natural), scaled by 1 / size
template <class Exec> void fr_ntt_inverse(Exec& ex, const FrNttPlan& p, const FrNttTables& tb, Fr* a, uint32_t total, uint32_t size) {
  FrNttPass pass[8];
  const int np = fr_ntt_passes(size, pass);
  for (int i = np - 1; i >= 0; i--) {
    if (ex.ntt_fused(true, a, total, pass[i].len, pass[i].k, (const Fr*)tb.twi, p.m)) continue;
    for (uint32_t s = pass[i].k; s-- > 0;) {
      uint32_t len = pass[i].len >> s;
      ex.template launch<FrNttDit>(total / 2, total / 2, len, p.m / len, (const Fr*)tb.twi, a);
    }
  }
  uint32_t shift = 0;
  while ((size << shift) < p.m) shift++;
  ex.template launch<FrScalePow2>(total, total, shift, (const Fr*)tb.consts, a);
}

// Once per n.  scratch: 5 m elements.
template <class Exec> void fr_quotient_setup(Exec& ex, const FrNttPlan& p, const FrNttTables& tb, Fr* scratch) {
  const uint32_t m = p.m, n2 = m / 2, n = p.n;
  ex.template launch<FrNttSetup>(1u, p.logm, tb.consts);
  ex.template launch<FrPowTable>(m / 2, m / 2, (const Fr*)tb.consts, tb.tw);
  ex.template launch<FrPowTable>(m / 2, m / 2, (const Fr*)(tb.consts + 1), tb.twi);
  Fr* poly = scratch;            // n2
  Fr* poly2 = scratch + n2;      // n2
  Fr* x = scratch + m;           // m  (= 2 n2)
  Fr* y = scratch + 2 * m;       // n2 (tree) / m (Newton)
  Fr* z = scratch + 3 * m;       // m
  Fr* cur = scratch + 4 * m;     // m: the inverse series so far
  // product tree
  ex.template launch<FrTreeLeaves>(n2, n2, n, poly);
  for (uint32_t d = 1; d < n2; d <<= 1) {
    ex.template launch<FrTreeExpand>(m, m, d, (const Fr*)poly, x);
    fr_ntt_forward(ex, p, tb, x, m, 2 * d);
    ex.template launch<FrTreeMul>(n2, n2, d, (const Fr*)x, y);
    fr_ntt_inverse(ex, p, tb, y, n2, 2 * d);
    ex.template launch<FrTreeCombine>(n2, n2, d, (const Fr*)y, (const Fr*)poly, poly2);
    Fr* tmp = poly; poly = poly2; poly2 = tmp;
  }
  ex.template launch<FrTreeToT>(m, m, n2, n, (const Fr*)poly, tb.t);
  ex.template launch<FrCopyPad>(m, m, m, (const Fr*)tb.t, tb.that);
  fr_ntt_forward(ex, p, tb, tb.that, m, m);
  // Newton: I_1 = 1; I_2k = I_k (2 - f I_k) mod x^2k, transforms of size 4k <= m; stop once k >= n - 1
  ex.template launch<FrRevT>(m, m, 1u, n, (const Fr*)tb.t, cur);     // f mod x = 1 = I_1 (t is monic)
  for (uint32_t k = 1; k + 1 < n; k <<= 1) {
    const uint32_t s = 4 * k;                                       // transform size
    ex.template launch<FrRevT>(s, s, 2 * k, n, (const Fr*)tb.t, x);  // f mod x^2k
    ex.template launch<FrCopyPad>(s, s, k, (const Fr*)cur, y);        // I_k
    fr_ntt_forward(ex, p, tb, x, s, s);
    fr_ntt_forward(ex, p, tb, y, s, s);
    ex.template launch<FrPointMul>(s, s, (const Fr*)x, (const Fr*)y, x);
    fr_ntt_inverse(ex, p, tb, x, s, s);                               // e = f I_k (exact: degree < 3k)
    ex.template launch<FrTwoMinus>(s, s, 2 * k, (const Fr*)x, z);     // 2 - e mod x^2k
    fr_ntt_forward(ex, p, tb, z, s, s);
    ex.template launch<FrPointMul>(s, s, (const Fr*)z, (const Fr*)y, z);
    fr_ntt_inverse(ex, p, tb, z, s, s);
    ex.template launch<FrCopyPad>(m, m, 2 * k, (const Fr*)z, cur);    // I_2k
  }
  ex.template launch<FrCopyPad>(m, m, n - 1, (const Fr*)cur, tb.ihat);  // I mod x^(n-1)
  fr_ntt_forward(ex, p, tb, tb.ihat, m, m);
}

// Per proof.  raw_*: canonical words (n x 8) on the device; scratch: 4 m elements; out: (n - 1) x 8 words; *flag |= 1
// when p != h t.
template <class Exec>
void fr_quotient_run(Exec& ex, const FrNttPlan& p, const FrNttTables& tb, const uint32_t* raw_u, const uint32_t* raw_v,
                     const uint32_t* raw_w, Fr* scratch, uint32_t* out, uint32_t* flag) {
  const uint32_t m = p.m, n = p.n;
  Fr* a = scratch;
  Fr* b = scratch + m;
  Fr* pp = scratch + 2 * m;
  Fr* h = scratch + 3 * m;
  ex.template launch<FrVecToMont>(m, m, n, raw_u, a);
  ex.template launch<FrVecToMont>(m, m, n, raw_v, b);
  fr_ntt_forward(ex, p, tb, a, m, m);
  fr_ntt_forward(ex, p, tb, b, m, m);
  ex.template launch<FrPointMul>(m, m, (const Fr*)a, (const Fr*)b, pp);
  fr_ntt_inverse(ex, p, tb, pp, m, m);
  ex.template launch<FrVecToMont>(n, n, n, raw_w, b);
  ex.template launch<FrSubW>(n, n, (const Fr*)b, pp);                       // p = u v - w
  ex.template launch<FrRevTop>(m, m, n, (const Fr*)pp, a);
  fr_ntt_forward(ex, p, tb, a, m, m);
  ex.template launch<FrPointMul>(m, m, (const Fr*)a, (const Fr*)tb.ihat, a);
  fr_ntt_inverse(ex, p, tb, a, m, m);                                        // rev(h) in the low n - 1 coefficients
  ex.template launch<FrExtractH>(m, m, n, (const Fr*)a, h, out);
  fr_ntt_forward(ex, p, tb, h, m, m);
  ex.template launch<FrPointMul>(m, m, (const Fr*)h, (const Fr*)tb.that, h);
  fr_ntt_inverse(ex, p, tb, h, m, m);                                        // h t
  ex.template launch<FrCheckEqual>(m, m, (const Fr*)h, (const Fr*)pp, flag);
}

}  // namespace zk
