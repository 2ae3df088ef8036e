// Block-cooperative point operations for the latency-bound tail of the MSM (bucket reduction, window tree).
//
// Measured on B200 (tools/ubench.cu, profiles/r1_ubench_latency.log): one warp issues an IMAD.WIDE every
// ~6 cycles whatever its instruction-level parallelism, so a lone warp needs 0.95 us per modular
// multiplication, 13.2 us per XYZZ addition and 8.4 us per doubling, and the ~45 dependent point operations of
// the reduction cost > 0.5 ms while most of the chip idles.  Here a block of 4 warps (one per scheduler of an
// SM) works on 32 chains at once (lane = chain): the independent field products of one point operation are
// spread over the 4 warps and meet in shared memory, so an addition is 4 multiplication rounds (instead of 14
// sequential products) and a doubling 3 (instead of 9).
//
// Shared-memory slot layout: slot s, limb j, lane l at word (s * NL + j) * 32 + l  (conflict-free).
// Same formulas and exceptional cases as ec.cuh (xyzz_add / xyzz_dbl); results are bit-identical.
#pragma once
#include "msm.cuh"

namespace zk {
namespace coop {

// Warp roles.  A block works on 32 chains (lane = chain) with 4 warp GROUPS, one per independent field product of a
// round.  For Fq a group is one warp.  An Fq2 product is three Fq products (Karatsuba: a0 b0, a1 b1, (a0 + a1)(b0 + b1)),
// so for G2 a group is THREE warps (12 per block, 3 per scheduler) that run them side by side: an Fq2 round then costs
// one Fq multiplication of latency plus a short join, not three multiplications one after the other.
template <class F> struct SubWarps { static constexpr int k = 1; };
template <> struct SubWarps<Fp2> { static constexpr int k = 3; };
template <class F> constexpr int block_threads() { return 4 * SubWarps<F>::k * 32; }
static constexpr int kMaxThreads = 384;   // launch bound of every cooperative kernel (G2's 12 warps)
template <class F> struct Role {
  int w, sub, l;   // group (0..3), warp within the group, lane
  __device__ __forceinline__ Role() {
    const int warp = threadIdx.x >> 5;
    w = warp / SubWarps<F>::k; sub = warp % SubWarps<F>::k; l = threadIdx.x & 31;
  }
  __device__ __forceinline__ bool lead() const { return sub == 0; }   // the warp that does the group's non-product work
};

template <class F> struct Slots {
  static constexpr int NL = sizeof(F) / 4;
  uint32_t* base;
  __device__ __forceinline__ F load(int slot, int lane) const {
    F r;
    const uint32_t* p = base + slot * NL * 32 + lane;
    uint32_t* w = reinterpret_cast<uint32_t*>(&r);
#pragma unroll
    for (int j = 0; j < NL; j++) w[j] = p[j * 32];
    return r;
  }
  __device__ __forceinline__ void store(int slot, int lane, const F& v) const {
    uint32_t* p = base + slot * NL * 32 + lane;
    const uint32_t* w = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
    for (int j = 0; j < NL; j++) p[j * 32] = w[j];
  }
  // The ONLY field multiplication of the cooperative kernels: a real call taking slot numbers (nothing is
  // passed through local memory), so each kernel holds one copy of the ~6 KB multiplication instead of ~30
  // and the four warps of a block, which run different products of a round, all fetch the same
  // instruction lines.  dst may alias a or b (both are loaded first).
  __device__ __noinline__ void mul(int dst, int a, int b, int lane) const {
    F x = load(a, lane), y = load(b, lane), r;
    fmul(r, x, y);
    store(dst, lane, r);
  }
  // ---- split products (SubWarps<F>::k == 3, i.e. F = Fq2): warp `sub` of group w computes one of the three Fq
  // products of a * b into the group's scratch (3 Fq slots per group behind the F slots), the lead warp joins them
  uint32_t* xbase;   // 12 Fq scratch slots: (w * 3 + sub), limb j, lane at ((w * 3 + sub) * 12 + j) * 32 + lane
  __device__ __forceinline__ Fp load_half(int slot, int half, int lane) const {
    Fp r;
    const uint32_t* p = base + slot * NL * 32 + half * 12 * 32 + lane;
#pragma unroll
    for (int j = 0; j < 12; j++) r.v[j] = p[j * 32];
    return r;
  }
  __device__ __noinline__ void mul_part(int w, int sub, int a, int b, int lane) const {
    Fp x, y, r;
    if (sub < 2) { x = load_half(a, sub, lane); y = load_half(b, sub, lane); }
    else {
      Fp a0 = load_half(a, 0, lane), a1 = load_half(a, 1, lane), b0 = load_half(b, 0, lane), b1 = load_half(b, 1, lane);
      fadd(x, a0, a1);
      fadd(y, b0, b1);
    }
    fmul(r, x, y);
    uint32_t* q = xbase + (w * 3 + sub) * 12 * 32 + lane;
#pragma unroll
    for (int j = 0; j < 12; j++) q[j * 32] = r.v[j];
  }
  __device__ __forceinline__ void mul_join(int w, int dst, int lane) const {   // c0 = t0 - t1, c1 = t2 - t0 - t1
    Fp t[3];
#pragma unroll
    for (int i = 0; i < 3; i++) {
      const uint32_t* q = xbase + (w * 3 + i) * 12 * 32 + lane;
#pragma unroll
      for (int j = 0; j < 12; j++) t[i].v[j] = q[j * 32];
    }
    Fp c0, c1;
    fsub(c0, t[0], t[1]);
    fsub(c1, t[2], t[0]);
    fsub(c1, c1, t[1]);
    uint32_t* d = base + dst * NL * 32 + lane;
#pragma unroll
    for (int j = 0; j < 12; j++) { d[j * 32] = c0.v[j]; d[(12 + j) * 32] = c1.v[j]; }
  }
};

// One multiplication round: group g computes dst[g] = a[g] * b[g] (dst[g] < 0: the group idles), then the block
// meets.  presync(): a round whose operands were just written by a lead warp needs the block to meet BEFORE the
// products when the other warps of the group read them (split products only).
template <class F> __device__ __forceinline__ void presync() {
  if (SubWarps<F>::k > 1) __syncthreads();
}
template <class F> __device__ __forceinline__ void mul_round(const Slots<F>& S, const Role<F>& r, int d0, int a0, int b0, int d1, int a1,
                                                             int b1, int d2 = -1, int a2 = 0, int b2 = 0, int d3 = -1, int a3 = 0,
                                                             int b3 = 0) {
  const int d = r.w == 0 ? d0 : (r.w == 1 ? d1 : (r.w == 2 ? d2 : d3));
  const int a = r.w == 0 ? a0 : (r.w == 1 ? a1 : (r.w == 2 ? a2 : a3));
  const int b = r.w == 0 ? b0 : (r.w == 1 ? b1 : (r.w == 2 ? b2 : b3));
  if constexpr (SubWarps<F>::k == 1) {
    if (d >= 0) S.mul(d, a, b, r.l);
    __syncthreads();
  } else {
    if (d >= 0) S.mul_part(r.w, r.sub, a, b, r.l);
    __syncthreads();
    if (d >= 0 && r.lead()) S.mul_join(r.w, d, r.l);
    __syncthreads();
  }
}

// point = 4 consecutive slots X, Y, ZZ, ZZZ
enum { PX = 0, PY = 1, PZZ = 2, PZZZ = 3 };
// temporaries used by add / dbl (relative to the temp base)
enum { T_U1 = 0, T_U2, T_S1, T_S2, T_P, T_R, T_PP, T_RR, T_ZZ12, T_ZZZ12, T_PPP, T_Q, T_TT, T_VV,
       T_RX, T_RY, T_RZZ, T_RZZZ, T_DX, T_DY, T_DZZ, T_DZZZ, T_COUNT };
// flag bytes per lane (in shared memory after the slots)
struct Flags { uint8_t zero_p[32], zero_r[32]; };

// A = 2 A in place.  Infinity (ZZ = 0) and Y = 0 come out as ZZ = 0 without special handling.
// (uses temporaries U1, U2, P, R, PP, RR, PPP, Q, TT, VV only -- never the R* / D* result slots of point_add)
template <class F> __device__ __noinline__ void point_dbl(Slots<F> S, int A, int T) {
  const Role<F> r;
  const int l = r.l;
  // round 1: V = (2Y)^2 | XX = X^2
  if (r.w == 0 && r.lead()) {
    F y = S.load(A + PY, l), u;
    fdbl(u, y);
    S.store(T + T_U1, l, u);                  // U
  }
  presync<F>();
  mul_round(S, r, T + T_PP, T + T_U1, T + T_U1,   // V
            T + T_RR, A + PX, A + PX);            // XX
  // round 2: W = U V | S = X V | M = 3 XX, MM = M^2 | ZZ' = V ZZ
  if (r.w == 2 && r.lead()) {
    F xx = S.load(T + T_RR, l), m;
    fdbl(m, xx);
    fadd(m, m, xx);
    S.store(T + T_P, l, m);                   // M
  }
  presync<F>();
  mul_round(S, r, T + T_PPP, T + T_U1, T + T_PP,  // W
            T + T_Q, A + PX, T + T_PP,            // S
            T + T_R, T + T_P, T + T_P,            // MM
            A + PZZ, T + T_PP, A + PZZ);
  // round 3: X3 = MM - 2S, TT = M (S - X3) | WY = W Y | ZZZ' = W ZZZ
  if (r.w == 0 && r.lead()) {
    F mm = S.load(T + T_R, l), s = S.load(T + T_Q, l), x3, t;
    fsub(x3, mm, s);
    fsub(x3, x3, s);
    fsub(t, s, x3);
    S.store(T + T_U2, l, x3);
    S.store(T + T_TT, l, t);
  }
  presync<F>();
  mul_round(S, r, T + T_TT, T + T_P, T + T_TT,
            T + T_VV, T + T_PPP, A + PY,
            A + PZZZ, T + T_PPP, A + PZZZ);
  // round 4: Y3 = TT - WY
  if (r.w == 0 && r.lead()) {
    F t = S.load(T + T_TT, l), wy = S.load(T + T_VV, l), y3;
    fsub(y3, t, wy);
    S.store(A + PY, l, y3);
    S.store(A + PX, l, S.load(T + T_U2, l));
  }
  __syncthreads();
}

// A += Q (both XYZZ, complete).  q_masked: this lane's Q counts as infinity (used for the windowed scalar mul).
template <class F> __device__ __noinline__ void point_add(Slots<F> S, Flags* fl, int A, int Q, int T, bool q_masked) {
  const Role<F> r;
  const int w = r.w, l = r.l;
  const bool inf_a = fis_zero(S.load(A + PZZ, l));
  const bool inf_q = q_masked || fis_zero(S.load(Q + PZZ, l));
  // round 1: U1 = X1 ZZ2 | U2 = X2 ZZ1 | S1 = Y1 ZZZ2 | S2 = Y2 ZZZ1
  mul_round(S, r, T + T_U1, A + PX, Q + PZZ,
            T + T_U2, Q + PX, A + PZZ,
            T + T_S1, A + PY, Q + PZZZ,
            T + T_S2, Q + PY, A + PZZZ);
  // round 2: P = U2 - U1, PP = P^2 | R = S2 - S1, RR = R^2 | ZZ1 ZZ2 | ZZZ1 ZZZ2
  if (w == 0 && r.lead()) {
    F u1 = S.load(T + T_U1, l), u2 = S.load(T + T_U2, l), p;
    fsub(p, u2, u1);
    S.store(T + T_P, l, p);
    fl->zero_p[l] = fis_zero(p);
  } else if (w == 1 && r.lead()) {
    F s1 = S.load(T + T_S1, l), s2 = S.load(T + T_S2, l), rr;
    fsub(rr, s2, s1);
    S.store(T + T_R, l, rr);
    fl->zero_r[l] = fis_zero(rr);
  }
  presync<F>();
  mul_round(S, r, T + T_PP, T + T_P, T + T_P,
            T + T_RR, T + T_R, T + T_R,
            T + T_ZZ12, A + PZZ, Q + PZZ,
            T + T_ZZZ12, A + PZZZ, Q + PZZZ);
  // round 3: PPP = P PP | Q = U1 PP | ZZ3 = ZZ12 PP
  mul_round(S, r, T + T_PPP, T + T_P, T + T_PP,
            T + T_Q, T + T_U1, T + T_PP,
            T + T_RZZ, T + T_ZZ12, T + T_PP);
  // round 4: X3 = RR - PPP - 2Q, TT = R (Q - X3) | VV = S1 PPP | ZZZ3 = ZZZ12 PPP
  if (w == 0 && r.lead()) {
    F rr = S.load(T + T_RR, l), ppp = S.load(T + T_PPP, l), q = S.load(T + T_Q, l), x3, t;
    fsub(x3, rr, ppp);
    fsub(x3, x3, q);
    fsub(x3, x3, q);
    fsub(t, q, x3);
    S.store(T + T_RX, l, x3);
    S.store(T + T_TT, l, t);
  }
  presync<F>();
  mul_round(S, r, T + T_TT, T + T_R, T + T_TT,
            T + T_VV, T + T_S1, T + T_PPP,
            T + T_RZZZ, T + T_ZZZ12, T + T_PPP);
  // round 5: Y3 = TT - VV
  if (w == 0 && r.lead()) {
    F t = S.load(T + T_TT, l), v = S.load(T + T_VV, l), y3;
    fsub(y3, t, v);
    S.store(T + T_RY, l, y3);
  }
  const bool zero_p = fl->zero_p[l] != 0, zero_r = fl->zero_r[l] != 0;
  const bool need_dbl = !inf_a && !inf_q && zero_p && zero_r;   // same point: tangent (macros.rs:57-108)
  if (__syncthreads_or(need_dbl)) {                              // rare; block-uniform branch
    if (r.lead()) S.store(T + T_DX + w, l, S.load(A + w, l));    // group w copies coordinate w
    __syncthreads();
    point_dbl(S, T + T_DX, T);                                   // temporaries U1..VV are free again
  }
  // select, group w handles coordinate w
  F res;
  if (r.lead()) {
    if (inf_q) res = S.load(A + w, l);
    else if (inf_a) res = S.load(Q + w, l);
    else if (zero_p) { if (zero_r) res = S.load(T + T_DX + w, l); else fset_zero(res); }   // P + (-P) = infinity
    else res = S.load(T + T_RX + w, l);
  }
  __syncthreads();
  if (r.lead()) S.store(A + w, l, res);
  __syncthreads();
}

template <class F> __device__ __forceinline__ void point_set_inf(const Slots<F>& S, int A) {
  const Role<F> r;
  F z;
  fset_zero(z);
  if (r.lead()) S.store(A + r.w, r.l, z);
}
template <class F> __device__ __forceinline__ void point_copy(const Slots<F>& S, int dst, int src) {
  const Role<F> r;
  if (r.lead()) S.store(dst + r.w, r.l, S.load(src + r.w, r.l));
}
// warp w moves coordinate w of a global XYZZ point to / from the lane's slot
template <class F> __device__ __forceinline__ void point_load_global(const Slots<F>& S, int A, const XYZZ<F>* g, bool present) {
  const Role<F> r;
  if (!r.lead()) return;
  F v;
  if (present) v = reinterpret_cast<const F*>(g)[r.w];
  else fset_zero(v);
  S.store(A + r.w, r.l, v);
}
template <class F> __device__ __forceinline__ void point_store_global(const Slots<F>& S, int A, XYZZ<F>* g) {
  const Role<F> r;
  if (r.lead()) reinterpret_cast<F*>(g)[r.w] = S.load(A + r.w, r.l);
}

enum { S_RUN = 0, S_ACC = 4, S_Q = 8, S_BASE = 12, S_B2 = 16, S_B3 = 20, S_TMP = 24, S_TOTAL = S_TMP + T_COUNT };

// shared memory of a block: the F slots, the split-product scratch (12 Fq slots, G2 only), the flag bytes
template <class F> __host__ __device__ constexpr size_t slot_words() { return (size_t)S_TOTAL * (sizeof(F) / 4) * 32; }
template <class F> __host__ __device__ constexpr size_t scratch_words() { return SubWarps<F>::k > 1 ? (size_t)12 * 12 * 32 : 0; }
template <class F> __device__ __forceinline__ Slots<F> make_slots(uint32_t* smem) { return Slots<F>{smem, smem + slot_words<F>()}; }
template <class F> __device__ __forceinline__ Flags* flags_of(uint32_t* smem) {
  return reinterpret_cast<Flags*>(smem + slot_words<F>() + scratch_words<F>());
}

// Sum over the lanes of a block: lane 0's S_ACC += the S_ACC of lanes 1 .. cnt-1 (log2 levels; lane l takes lane
// l + stride's point as its addend, only the lower half accumulates).  cnt is block-uniform; the caller has synced.
template <class F> __device__ __forceinline__ void lane_tree(const Slots<F>& S, Flags* fl, uint32_t cnt) {
  const Role<F> r;
  const int w = r.w, l = r.l;
  for (int stride = 16; stride >= 1; stride >>= 1) {
    if ((uint32_t)stride >= cnt) continue;
    if (r.lead()) {
      F v = S.load(S_ACC + w, (l + stride) & 31);
      S.store(S_Q + w, l, v);
    }
    __syncthreads();
    point_add(S, fl, S_ACC, S_Q, S_TMP, l >= stride);
  }
}

template <class F> constexpr size_t smem_bytes() { return (slot_words<F>() + scratch_words<F>()) * 4 + sizeof(Flags); }

// 32 chains per block; chain (win, k): out = sum_{i<K} (g + i + 1) * bucket[win*B + k*K + i], g = global index of the
// chain's first bucket (as BucketReduce)
template <class C>
__global__ void __launch_bounds__(kMaxThreads) bucket_reduce_kernel(MsmPlan p, const uint32_t* offsets,
                                                                 const XYZZ<typename C::F>* bucket_sums,
                                                                 XYZZ<typename C::F>* out, int tree) {
  typedef typename C::F F;
  extern __shared__ __align__(16) uint32_t smem[];
  Slots<F> S = make_slots<F>(smem);
  Flags* fl = flags_of<F>(smem);
  const int l = threadIdx.x & 31;
  const uint32_t chunks = p.B / p.K, total = p.nwin * chunks;
  const uint32_t chain = blockIdx.x * 32 + l;
  const bool valid = chain < total;
  const uint32_t win = valid ? chain / chunks : 0, k = valid ? chain % chunks : 0;
  const size_t base = (size_t)win * p.B + (size_t)k * p.K;
  // last bucket of the chain: run = acc = bucket (nothing is added to infinity)
  {
    const int i = (int)p.K - 1;
    bool present = valid && offsets[base + i] != offsets[base + i + 1];
    point_load_global(S, S_RUN, bucket_sums + base + i, present);
    __syncthreads();
    point_copy(S, S_ACC, S_RUN);
    __syncthreads();
  }
  for (int i = (int)p.K - 2; i >= 0; i--) {
    bool present = valid && offsets[base + i] != offsets[base + i + 1];
    point_load_global(S, S_Q, bucket_sums + base + i, present);
    __syncthreads();
    point_add(S, fl, S_RUN, S_Q, S_TMP, false);
    point_add(S, fl, S_ACC, S_RUN, S_TMP, false);
  }
  // run = g * run, g = global index of the chain's first bucket (msm_global_bucket: the chain's own k K without the
  // bucket-range split), over the bit length of the largest multiplier of any rank (uniform for the grid), MSB first
  // in 2-bit windows over the multiples 1x, 2x, 3x of run: half the additions of the bit-by-bit method.  g is a
  // multiple of K, so the low log2(K) bits are zero: doublings only.
  const uint32_t s = msm_global_bucket(p, k * p.K);
  const int nbits = 32 - __clz((((uint32_t)p.B << p.world_log) - p.K) | 1u);
  const int low = __ffs((int)p.K) - 1;
  const Role<F> role;
  point_copy(S, S_BASE, S_RUN);
  point_copy(S, S_B2, S_RUN);
  __syncthreads();
  point_dbl(S, S_B2, S_TMP);
  point_copy(S, S_B3, S_B2);
  __syncthreads();
  point_add(S, fl, S_B3, S_BASE, S_TMP, false);
  auto window_to = [&](int dst, uint32_t w) {   // lane's dst = w * base (w in 0..3; 0 = infinity); group g moves coordinate g
    if (!role.lead()) return;
    F v;
    if (w == 0) fset_zero(v);
    else v = S.load((w == 1 ? S_BASE : (w == 2 ? S_B2 : S_B3)) + role.w, l);
    S.store(dst + role.w, l, v);
  };
  int b = nbits;
  if (b <= low) {
    point_set_inf(S, S_RUN);   // multiplier 0 for every chain (a single chunk)
    __syncthreads();
  } else {
    const int top = ((b - low) & 1) ? 1 : 2;   // odd number of positions: the top window is one bit
    b -= top;
    window_to(S_RUN, (s >> b) & (top == 1 ? 1u : 3u));
    __syncthreads();
    while (b > low) {
      b -= 2;
      point_dbl(S, S_RUN, S_TMP);
      point_dbl(S, S_RUN, S_TMP);
      const uint32_t w = (s >> b) & 3u;
      window_to(S_Q, w);
      __syncthreads();
      point_add(S, fl, S_RUN, S_Q, S_TMP, w == 0);
    }
    for (int i = 0; i < low; i++) point_dbl(S, S_RUN, S_TMP);
  }
  point_add(S, fl, S_ACC, S_RUN, S_TMP, false);
  if (!tree) {
    if (valid) point_store_global(S, S_ACC, out + chain);
    return;
  }
  // chunks % 32 == 0: the block's 32 chains belong to one window; their sum is element k / 32 of that window's row
  lane_tree(S, fl, 32);
  if (l == 0 && valid) point_store_global(S, S_ACC, out + (size_t)win * chunks + k / 32);
}

// One level of the window tree: block (win, j) sums elements [j per_block, (j + 1) per_block) of row win of `in`
// (row length m) into element j of row win of `out` (a different buffer).  per_block / 32 - 1 additions per lane,
// then the lane tree.
template <class C>
__global__ void __launch_bounds__(kMaxThreads) row_sum_kernel(uint32_t pitch_in, uint32_t m, uint32_t per_block,
                                                           uint32_t blocks_per_row, const XYZZ<typename C::F>* in,
                                                           uint32_t pitch_out, XYZZ<typename C::F>* out) {
  typedef typename C::F F;
  extern __shared__ __align__(16) uint32_t smem[];
  Slots<F> S = make_slots<F>(smem);
  Flags* fl = flags_of<F>(smem);
  const uint32_t l = threadIdx.x & 31;
  const uint32_t win = blockIdx.x / blocks_per_row, j = blockIdx.x % blocks_per_row;
  const uint32_t first = j * per_block;
  const uint32_t cnt = m - first < per_block ? m - first : per_block;
  const XYZZ<F>* row = in + (size_t)win * pitch_in + first;
  point_load_global(S, S_ACC, row + l, l < cnt);
  __syncthreads();
  for (uint32_t t = 32; t < cnt; t += 32) {
    point_load_global(S, S_Q, row + t + l, t + l < cnt);
    __syncthreads();
    point_add(S, fl, S_ACC, S_Q, S_TMP, false);
  }
  lane_tree(S, fl, cnt);
  if (l == 0) point_store_global(S, S_ACC, out + (size_t)win * pitch_out + j);
}

// lane 0 of a single block: Horner over the window sums (row w of arr, element 0), then the outputs of Finish
// (msm.cuh): the XYZZ partial and / or the canonical affine point.  c doublings per window at ~5 us instead of 8.4.
template <class C>
__global__ void __launch_bounds__(kMaxThreads) finish_kernel(uint32_t nwin, uint32_t pitch, uint32_t c,
                                                          const XYZZ<typename C::F>* arr, XYZZ<typename C::F>* out_xyzz,
                                                          uint32_t* out_affine, uint32_t* out_inf) {
  typedef typename C::F F;
  extern __shared__ __align__(16) uint32_t smem[];
  Slots<F> S = make_slots<F>(smem);
  Flags* fl = flags_of<F>(smem);
  const int l = threadIdx.x & 31;
  point_load_global(S, S_ACC, arr + (size_t)(nwin - 1) * pitch, l == 0);
  __syncthreads();
  for (int w = (int)nwin - 2; w >= 0; w--) {
    for (uint32_t i = 0; i < c; i++) point_dbl(S, S_ACC, S_TMP);
    point_load_global(S, S_Q, arr + (size_t)w * pitch, l == 0);
    __syncthreads();
    point_add(S, fl, S_ACC, S_Q, S_TMP, false);
  }
  if (threadIdx.x == 0) {
    XYZZ<F> acc;
    acc.x = S.load(S_ACC + PX, 0); acc.y = S.load(S_ACC + PY, 0);
    acc.zz = S.load(S_ACC + PZZ, 0); acc.zzz = S.load(S_ACC + PZZZ, 0);
    if (out_xyzz) *out_xyzz = acc;
    if (out_affine) {
      Affine<F> a;
      xyzz_to_affine(a, acc);
      store_canonical<F>(out_affine, a);
      *out_inf = is_inf(acc) ? 1u : 0u;
    }
  }
}

// sum of k <= 32 partial points (multi-GPU combine): lane i holds partial i, 5-level tree across lanes
template <class C>
__global__ void __launch_bounds__(kMaxThreads) combine_kernel(uint32_t k, const XYZZ<typename C::F>* parts, uint32_t stride_words,
                                                           uint32_t* out_affine, uint32_t* out_inf, uint32_t* err) {
  typedef typename C::F F;
  extern __shared__ __align__(16) uint32_t smem[];
  Slots<F> S = make_slots<F>(smem);
  Flags* fl = flags_of<F>(smem);
  const int l = threadIdx.x & 31;
  bool present = (uint32_t)l < k;
  // partial l sits stride_words 32-bit words after partial l - 1 (packed arrays: sizeof(XYZZ) / 4; the per-rank
  // blobs of a distributed proof: ZKMSM_GROTH16_PARTIAL_WORDS)
  const XYZZ<F>* mine = reinterpret_cast<const XYZZ<F>*>(reinterpret_cast<const uint32_t*>(parts) + (size_t)(present ? l : 0) * stride_words);
  if (present && is_poisoned(*mine)) {   // a rank saw an out-of-range scalar (PoisonPartial, msm.cuh)
    if (threadIdx.x < 32) atomicOr(err, ERR_SCALAR_RANGE);
    present = false;
  }
  point_load_global(S, S_ACC, mine, present);
  __syncthreads();
  lane_tree(S, fl, k);
  if (threadIdx.x == 0) {
    XYZZ<F> acc;
    acc.x = S.load(S_ACC + PX, 0); acc.y = S.load(S_ACC + PY, 0);
    acc.zz = S.load(S_ACC + PZZ, 0); acc.zzz = S.load(S_ACC + PZZZ, 0);
    Affine<F> a;
    xyzz_to_affine(a, acc);
    store_canonical<F>(out_affine, a);
    *out_inf = is_inf(acc) ? 1u : 0u;
  }
}

}  // namespace coop
}  // namespace zk
