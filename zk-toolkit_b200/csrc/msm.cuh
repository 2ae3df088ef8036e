// Pippenger multi-scalar multiplication  sum_i s_i * P_i  for sm_100a.
//
// Replaces the reference's serial loop `sum = sum + (&powers[i] * &self.coeffs[i])`
// (Polynomial::eval_with_g1_hidings / eval_with_g2_hidings,
// src/building_block/field/polynomial.rs:272-293) and the LSB-first affine double-and-add
// it calls per term (impl_scalar_mul_point!, src/building_block/curves/macros.rs:2-32).
//
// Pipeline (every stage is a kernel "body": one logical thread = one call of Body::run):
//   1 RecodeCount   scalar -> W signed c-bit digits; histogram of bucket sizes (atomics)
//   2 Scan*         exclusive scan of the histogram -> bucket offsets, scatter cursors
//   3 Scatter       counting sort: (bucket, +-point index) pairs grouped by bucket
//   3b BatchedAddRound  (large G1 sets) R rounds of pairwise affine additions per bucket, the inversions of T
//                   additions shared by Montgomery's trick; 2^R times fewer items reach stage 4
//   4 Accumulate    fixed-size chunks of L sorted pairs per thread, XYZZ mixed adds; a bucket
//                   that spans several chunks leaves one head sum + per-chunk partial sums
//   5 FixupLevel    16-ary tree over the per-chunk partial sums -> complete bucket sums
//   6 BucketReduce  per K consecutive buckets: sum_j (j+1) * bucket_j by running sums
//   7 window tree   chunk results of each window -> window sums (coop::row_sum_kernel; PairSum for wide levels)
//   8 Finish        Horner over windows (c doublings each), to canonical affine
//
// Work is independent of the scalar distribution: stage 4 always runs ceil(total/L) threads of
// <= L mixed adds, so all-equal scalars or duplicate points cost no more than random ones
// (the group law in ec.cuh is complete, so P+P and P+(-P) inside a bucket are exact).
//
// "Precomputed" point sets (built once when the CRS is loaded) hold 2^(c w) P_i for every
// window w; then all windows share ONE bucket set and stage 8 needs no doublings.
//
// The bodies are plain functions of (tid, args) so that tests/host_emu can run the identical
// pipeline on the CPU (test infrastructure only); the product launches them as CUDA kernels.
#pragma once
#include <stdlib.h>
#include "ec.cuh"

namespace zk {

#if defined(__CUDA_ARCH__)
ZK_D uint32_t zk_atomic_add(uint32_t* p, uint32_t v) { return atomicAdd(p, v); }
ZK_D void zk_atomic_or(uint32_t* p, uint32_t v) { atomicOr(p, v); }
// pull the cache lines of [p, p + bytes) towards the SM ahead of use (no registers held)
ZK_D void zk_prefetch(const void* p, uint32_t bytes) {
  const char* c = (const char*)p;
  asm volatile("prefetch.global.L1 [%0];" ::"l"(c));
  if ((((uintptr_t)c) & 127u) + bytes > 128u) asm volatile("prefetch.global.L1 [%0];" ::"l"(c + bytes - 1));
  if (bytes > 128u) asm volatile("prefetch.global.L1 [%0];" ::"l"(c + 128));
}
// start the DRAM -> L2 fetch of [p, p + bytes) two iterations ahead of a gather (holds no registers)
ZK_D void zk_prefetch_l2(const void* p, uint32_t bytes) {
  const char* c = (const char*)p;
#pragma unroll
  for (uint32_t off = 0; off < bytes; off += 64) asm volatile("prefetch.global.L2 [%0];" ::"l"(c + off));
  asm volatile("prefetch.global.L2 [%0];" ::"l"(c + bytes - 1));
}
#else
inline void zk_prefetch_l2(const void*, uint32_t) {}
inline void zk_prefetch(const void*, uint32_t) {}
inline uint32_t zk_atomic_add(uint32_t* p, uint32_t v) { uint32_t o = *p; *p = o + v; return o; }
inline void zk_atomic_or(uint32_t* p, uint32_t v) { *p |= v; }
#endif

struct G1 { typedef Fp F; static constexpr int AFF_LIMBS = 24; };
struct G2 { typedef Fp2 F; static constexpr int AFF_LIMBS = 48; };

static constexpr int SCALAR_LIMBS = 8;
static constexpr uint32_t ERR_SCALAR_RANGE = 1u;  // a scalar had bit 255 set (or was >= r for a subgroup point set)

struct MsmPlan {
  uint32_t n;          // terms
  uint32_t c;          // window bits
  uint32_t W;          // windows = floor(255 / c) + 1
  uint32_t a;          // windows 0 .. a-1 are c bits wide, windows a .. W-1 are c-1 bits wide (a = W: all c bits, the top
                       // one running out of scalar; precomputed sets balance the widths, see msm_wide_windows)
  uint32_t B;          // buckets per window = 2^(c-1)
  uint32_t nwin;       // bucket sets: W, or 1 with precomputed points
  uint32_t nb;         // total buckets = nwin * B
  uint32_t precomp;    // 1: points array holds W slabs of `stride` points, slab w = 2^(c w) P
  uint32_t stride;     // points per slab
  uint32_t L;          // sorted pairs per accumulate thread
  uint32_t K;          // buckets per reduce thread (power of two, divides B)
  uint32_t max_entries;   // n * W
  uint32_t acc_threads;   // ceil(max_entries / L)
  uint32_t batch_rounds;  // > 0: that many rounds of pairwise batched-affine additions before the XYZZ accumulation
  uint32_t batch_T;       //      additions per thread and inversion in those rounds
  uint32_t coop;          // 1: stages 6/7 run as block-cooperative kernels (coop.cuh), K sized for ~one block per SM
  uint32_t acc_slots;     // accumulate threads the device keeps resident at 2 blocks of 128 per SM (0 = unknown)
  uint32_t half;          // 1: point set is in the prime-order subgroup; scalars s > (r-1)/2 become r - s with
                          //    the point negated, so 254 bits are recoded and no carry-only top window exists
  // bucket-range split (precomputed sets) over 2^world_log ranks: the 2^(c-1) global buckets are cut into stripes of
  // 2^stripe_log consecutive buckets dealt out round-robin, this plan owns the stripes = rank (mod world), B of the
  // buckets in all, and keeps only the digits that fall there (striped, not contiguous: the balanced windows of a
  // precomputed set load the lower half of the bucket range more than the upper one).  world_log = 0: everything.
  uint32_t rank, world_log, stripe_log;
  uint32_t acc_G;         // > 0: bucket sums by AccumulateBuckets with acc_G lanes per bucket (no fix-up tree unless
                          //    a bucket holds more than acc_cap items); 0: chunked Accumulate + fix-up tree
  uint32_t acc_cap;       //    items per bucket above which the chunked fallback takes over
};

// global bucket (0-based magnitude - 1) -> this rank's local bucket, or false when another rank owns it
ZK_HD bool msm_local_bucket(const MsmPlan& p, uint32_t mag, uint32_t* local) {
  const uint32_t stripe = mag >> p.stripe_log;
  if ((stripe & ((1u << p.world_log) - 1u)) != p.rank) return false;
  *local = ((stripe >> p.world_log) << p.stripe_log) | (mag & ((1u << p.stripe_log) - 1u));
  return true;
}
// local bucket -> global bucket
ZK_HD uint32_t msm_global_bucket(const MsmPlan& p, uint32_t local) {
  const uint32_t stripe = local >> p.stripe_log;
  return (((stripe << p.world_log) | p.rank) << p.stripe_log) | (local & ((1u << p.stripe_log) - 1u));
}

// ---------------------------------------------------------------- signed-digit recoding
// Window w covers `width(w)` bits from `pos(w)`: the first `a` windows are c bits wide, the others c - 1 (a = W: the
// plain layout, every window c bits, the top one short because the scalar ends).  Digits d_w of a window of width
// t lie in [-(2^(t-1) - 1), 2^(t-1)]; the LAST window is not recoded (nothing could take its carry): it holds its
// bits plus the incoming carry, at most 2^(c-1), still a valid bucket.  sum_w d_w 2^pos(w) = scalar.
struct Digits {
  const uint32_t* s;
  uint32_t c, a, W, carry;
  ZK_HD Digits(const uint32_t* scalar, uint32_t c_, uint32_t a_, uint32_t W_) : s(scalar), c(c_), a(a_), W(W_), carry(0) {}
  ZK_HD int32_t next(uint32_t w) {
    const uint32_t t = w < a ? c : c - 1;                             // width
    const uint32_t bit = w < a ? w * c : a * c + (w - a) * (c - 1);   // position
    uint32_t limb = bit >> 5, sh = bit & 31;
    uint32_t v = 0;
    if (limb < SCALAR_LIMBS) {
      v = s[limb] >> sh;
      if (sh + t > 32 && limb + 1 < SCALAR_LIMBS) v |= s[limb + 1] << (32 - sh);
    }
    v = (v & ((1u << t) - 1)) + carry;
    if (w + 1 < W && v > (1u << (t - 1))) { carry = 1; return (int32_t)v - (int32_t)(1u << t); }
    carry = 0;
    return (int32_t)v;
  }
};

// Loads scalar `tid`.  Returns false (scalar rejected) when it is out of range.  With p.half the scalar must
// be < r; if it exceeds (r-1)/2 it is replaced by r - s and *negate is set: (r - s)(-P) = sP because the
// points of such a set have order r.
ZK_HD bool load_scalar(const MsmPlan& p, const uint32_t* scalars, uint32_t tid, uint32_t* s, bool* negate) {
#pragma unroll
  for (int i = 0; i < SCALAR_LIMBS; i++) s[i] = scalars[(size_t)tid * SCALAR_LIMBS + i];
  *negate = false;
  if (!p.half) return (s[SCALAR_LIMBS - 1] >> 31) == 0;
  uint32_t d[SCALAR_LIMBS];
  d[0] = ptx::sub_cc(FR_P[0], s[0]);                      // d = r - s
#pragma unroll
  for (int i = 1; i < SCALAR_LIMBS; i++) d[i] = ptx::subc_cc(FR_P[i], s[i]);
  uint32_t borrow = ptx::subc(0, 0);
  bool is_zero = true;
#pragma unroll
  for (int i = 0; i < SCALAR_LIMBS; i++) is_zero = is_zero && d[i] == 0;
  if (borrow || is_zero) return false;                    // s >= r
  uint32_t t = ptx::sub_cc(FR_HALF[0], s[0]);             // (r-1)/2 - s < 0  <=>  s in the upper half
#pragma unroll
  for (int i = 1; i < SCALAR_LIMBS; i++) t = ptx::subc_cc(FR_HALF[i], s[i]);
  (void)t;
  if (ptx::subc(0, 0)) {
    *negate = true;
#pragma unroll
    for (int i = 0; i < SCALAR_LIMBS; i++) s[i] = d[i];
  }
  return true;
}

struct RecodeCount {
  static const char* name() { return "recode_count"; }
  static ZK_HD void run(uint32_t tid, MsmPlan p, const uint32_t* scalars, uint32_t* hist, uint32_t* err) {
    if (tid >= p.n) return;
    uint32_t s[SCALAR_LIMBS];
    bool negate;
    if (!load_scalar(p, scalars, tid, s, &negate)) { zk_atomic_or(err, ERR_SCALAR_RANGE); return; }
    Digits dg(s, p.c, p.a, p.W);
    for (uint32_t w = 0; w < p.W; w++) {
      int32_t d = dg.next(w);
      if (d == 0) continue;
      uint32_t mag;
      if (!msm_local_bucket(p, (uint32_t)(d < 0 ? -d : d) - 1, &mag)) continue;   // another rank's bucket (range split)
      zk_atomic_add(&hist[(p.precomp ? 0 : w * p.B) + mag], 1u);
    }
  }
};

// ---------------------------------------------------------------- exclusive scan
// Reference formulation as three bodies (used by the CPU emulation of the pipeline); the CUDA build
// runs the block-cooperative kernels of tu_sort.cu instead (CudaExec::exclusive_scan), same result.
static constexpr uint32_t SCAN_SEG = 1024;  // elements per scan block; segsum holds nb / SCAN_SEG + 1 sums

struct ScanLocal {  static const char* name() { return "scan_local"; }  // thread t: sum of hist[t*SEG .. (t+1)*SEG)
  static ZK_HD void run(uint32_t tid, uint32_t nb, const uint32_t* hist, uint32_t* segsum) {
    uint32_t beg = tid * SCAN_SEG;
    if (beg >= nb) return;
    uint32_t end = beg + SCAN_SEG < nb ? beg + SCAN_SEG : nb, s = 0;
    for (uint32_t i = beg; i < end; i++) s += hist[i];
    segsum[tid] = s;
  }
};
struct ScanTop {  static const char* name() { return "scan_top"; }  // one thread: exclusive scan of the segment sums; grand total -> offsets[nb]
  static ZK_HD void run(uint32_t tid, uint32_t nseg, uint32_t nb, uint32_t* segsum, uint32_t* offsets) {
    if (tid != 0) return;
    uint32_t run = 0;
    for (uint32_t i = 0; i < nseg; i++) { uint32_t v = segsum[i]; segsum[i] = run; run += v; }
    offsets[nb] = run;
  }
};
struct ScanApply {  static const char* name() { return "scan_apply"; }  // offsets[] and the scatter cursors (cursor aliases hist)
  static ZK_HD void run(uint32_t tid, uint32_t nb, uint32_t* hist_cursor, const uint32_t* segsum, uint32_t* offsets) {
    uint32_t beg = tid * SCAN_SEG;
    if (beg >= nb) return;
    uint32_t end = beg + SCAN_SEG < nb ? beg + SCAN_SEG : nb, run = segsum[tid];
    for (uint32_t i = beg; i < end; i++) { uint32_t v = hist_cursor[i]; offsets[i] = run; hist_cursor[i] = run; run += v; }
  }
};

// ---------------------------------------------------------------- counting-sort scatter
struct alignas(8) Entry { uint32_t key, val; };  // val = point index | sign << 31

struct Scatter {
  static const char* name() { return "scatter"; }
  static ZK_HD void run(uint32_t tid, MsmPlan p, const uint32_t* scalars, uint32_t* cursor, Entry* entries) {
    if (tid >= p.n) return;
    uint32_t s[SCALAR_LIMBS];
    bool negate;
    if (!load_scalar(p, scalars, tid, s, &negate)) return;
    Digits dg(s, p.c, p.a, p.W);
    for (uint32_t w = 0; w < p.W; w++) {
      int32_t d = dg.next(w);
      if (d == 0) continue;
      uint32_t neg = d < 0, mag;
      if (!msm_local_bucket(p, (uint32_t)(neg ? -d : d) - 1, &mag)) continue;
      neg ^= negate ? 1u : 0u;
      uint32_t key = (p.precomp ? 0 : w * p.B) + mag;
      uint32_t idx = p.precomp ? w * p.stride + tid : tid;
      uint32_t pos = zk_atomic_add(&cursor[key], 1u);
      Entry e; e.key = key; e.val = idx | (neg << 31);
      entries[pos] = e;
    }
  }
};

// ---------------------------------------------------------------- bucket accumulation
static constexpr uint32_t NO_KEY = 0xffffffffu;
// partials folded per fix-up thread.  Level 0 is wide and throughput-bound (fan 4); the next two levels still carry
// work for ordinary buckets and are latency-bound at ~13 us per addition (fan 2); above that only giant buckets
// (skewed scalars) have anything left, so the tree closes quickly (fan 16).
inline uint32_t fix_fan(uint32_t level) { return level == 0 ? 4u : (level <= 2 ? 2u : 16u); }
// the same tree as the gated fallback of AccumulateBuckets / FixupDirect: it runs for skewed inputs only, so fewer,
// wider levels (every launch that finds the gate closed still costs a few microseconds)
inline uint32_t fix_fan_fallback(uint32_t) { return 32u; }

template <class C> struct Accumulate {
  typedef typename C::F F;
  static const char* name() { return "accumulate"; }
  // gate (nullable): the launch is the fallback of AccumulateBuckets and does nothing unless *gate != 0
  static ZK_HD void run(uint32_t tid, MsmPlan p, const uint32_t* offsets, const Entry* entries,
                        const Affine<F>* points, XYZZ<F>* bucket_sums, XYZZ<F>* partials, uint32_t* partial_keys,
                        const uint32_t* gate) {
    if (gate && *gate == 0) return;
    uint32_t total = offsets[p.nb];
    uint64_t beg64 = (uint64_t)tid * p.L;
    if (beg64 >= total) { partial_keys[tid] = NO_KEY; return; }
    uint32_t beg = (uint32_t)beg64, end = beg + p.L < total ? beg + p.L : total;
    uint32_t key = entries[beg].key;
    bool head = beg == 0 || entries[beg - 1].key != key;   // segment starts its bucket?
    partial_keys[tid] = head ? NO_KEY : key;
    XYZZ<F> acc;
    set_inf(acc);
    Entry e = entries[beg];
#if defined(ZK_ACC_DOUBLE_BUFFER)
    // software pipeline: the next point is loaded into registers before the current mixed add starts
    Affine<F> qn = points[e.val & 0x7fffffffu];
    for (uint32_t pos = beg; pos < end; pos++) {
      Entry en = pos + 1 < end ? entries[pos + 1] : e;
      Affine<F> q = qn;
      qn = points[en.val & 0x7fffffffu];
      if (e.key != key) {
        if (head) bucket_sums[key] = acc; else partials[tid] = acc;
        set_inf(acc);
        key = e.key;
        head = true;
      }
      affine_cneg(q, (e.val >> 31) != 0);
      xyzz_madd(acc, q);
      e = en;
    }
#else
    for (uint32_t pos = beg; pos < end; pos++) {
      Entry en = pos + 1 < end ? entries[pos + 1] : e;
      zk_prefetch(&points[en.val & 0x7fffffffu], sizeof(Affine<F>));   // next point rides under this mixed add
      if (e.key != key) {
        if (head) bucket_sums[key] = acc; else partials[tid] = acc;
        set_inf(acc);
        key = e.key;
        head = true;
      }
      Affine<F> q = points[e.val & 0x7fffffffu];
      affine_cneg(q, (e.val >> 31) != 0);
      xyzz_madd(acc, q);
      e = en;
    }
#endif
    if (head) bucket_sums[key] = acc; else partials[tid] = acc;
  }
};

// Bucket sums with G lanes per bucket (G a power of two <= 32, chosen so that the launch fills the device): lane g
// adds the items s + g, s + g + G, .. of its bucket with mixed additions; the G partial sums then meet in a
// log2(G)-level tree (register shuffles on the device, BucketAccTree on the CPU emulation).  Every bucket is
// complete after this one launch -- no partial sums, no fix-up tree -- which is what the batched-affine rounds
// leave behind (8..24 items per bucket) and what small shards look like.  A bucket with more than p.acc_cap items
// (skewed scalars) raises *big and is left alone: the chunked Accumulate + FixupLevel launches that follow are
// gated on that flag and then redo all buckets, balanced for any distribution.
//   direct != 0: item i is points[i] (output of a batched round); else item i is entries[i] -> (index, sign).
template <class C> struct BucketAccLane {
  typedef typename C::F F;
  static ZK_HD Affine<F> item(uint32_t i, const Entry* entries, const Affine<F>* points, uint32_t direct) {
    if (direct) return points[i];
    Entry e = entries[i];
    Affine<F> q = points[e.val & 0x7fffffffu];
    affine_cneg(q, (e.val >> 31) != 0);
    return q;
  }
  // returns false when the bucket is over the cap (acc is infinity then)
  static ZK_HD bool run(XYZZ<F>& acc, uint32_t s, uint32_t e, uint32_t g, uint32_t G, uint32_t cap, const Entry* entries,
                        const Affine<F>* points, uint32_t direct) {
    set_inf(acc);
    if (e - s > cap) return false;
    uint32_t i = s + g;
    if (i >= e) return true;
    if (sizeof(F) > 48) {
      // G2: accumulator (96 words) + one point (48) already fill the register file; the next point is only pulled
      // towards the SM (no registers held) instead of being double-buffered
      for (; i < e; i += G) {
        if (i + G < e) {
          if (direct) zk_prefetch(&points[i + G], sizeof(Affine<F>));
          else zk_prefetch(&points[entries[i + G].val & 0x7fffffffu], sizeof(Affine<F>));
        }
        Affine<F> q = item(i, entries, points, direct);
        xyzz_madd(acc, q);
      }
      return true;
    }
    Affine<F> qn = item(i, entries, points, direct);
    for (; i < e; i += G) {
      Affine<F> q = qn;
      if (i + G < e) qn = item(i + G, entries, points, direct);   // next item rides under this mixed add
      xyzz_madd(acc, q);
    }
    return true;
  }
};
// second half of the G2 form of the launch: thread b adds the acc_G lane sums of bucket b (left in lane_sums[b G ..])
template <class C> struct BucketLaneSum {
  typedef typename C::F F;
  static const char* name() { return "bucket_lane_sum"; }
  static ZK_HD void run(uint32_t tid, MsmPlan p, const uint32_t* offsets, const XYZZ<F>* lane_sums, XYZZ<F>* bucket_sums,
                        const uint32_t* big) {
    if (tid >= p.nb || offsets[tid] == offsets[tid + 1] || *big) return;   // (a raised flag: the fallback redoes every bucket)
    XYZZ<F> acc = lane_sums[(size_t)tid * p.acc_G];
    for (uint32_t g = 1; g < p.acc_G; g++) { XYZZ<F> q = lane_sums[(size_t)tid * p.acc_G + g]; xyzz_add(acc, q); }
    bucket_sums[tid] = acc;
  }
};
// the emulation's form of the launch: one logical thread per bucket runs the G lanes one after the other and
// folds them in the order of the device's shuffle tree
template <class C> struct AccumulateBucketsRef {
  typedef typename C::F F;
  static const char* name() { return "accumulate_buckets"; }
  static ZK_HD void run(uint32_t tid, MsmPlan p, const uint32_t* offsets, const Entry* entries, const Affine<F>* points,
                        uint32_t direct, XYZZ<F>* bucket_sums, uint32_t* big) {
    if (tid >= p.nb) return;
    const uint32_t s = offsets[tid], e = offsets[tid + 1], G = p.acc_G;
    if (s == e) return;
    XYZZ<F> lane[32];
    bool ok = true;
    for (uint32_t g = 0; g < G; g++) ok = BucketAccLane<C>::run(lane[g], s, e, g, G, p.acc_cap, entries, points, direct) && ok;
    if (!ok) { zk_atomic_or(big, 1u); return; }
    for (uint32_t d = G / 2; d >= 1; d >>= 1)
      for (uint32_t g = 0; g < d; g++) xyzz_add(lane[g], lane[g + d]);
    bucket_sums[tid] = lane[0];
  }
};

// One level of the fix-up tree.  Input: `count` partial sums in bucket order, keys_in[t] = bucket the
// partial continues (NO_KEY = none).  Thread u folds `fan` consecutive partials: a run of equal keys
// that STARTS inside the thread's range is added to its bucket sum (exactly one thread per bucket and
// level does that); a run that continues from the previous range is passed on as one partial of the
// next level.  Balanced for any scalar distribution: a bucket spanning m chunks costs log_16(m) levels.
template <class C> struct FixupLevel {
  typedef typename C::F F;
  static const char* name() { return "fixup_level"; }
  static ZK_HD void run(uint32_t tid, uint32_t count, uint32_t fan, const uint32_t* keys_in, const XYZZ<F>* parts_in,
                        uint32_t* keys_out, XYZZ<F>* parts_out, XYZZ<F>* bucket_sums, const uint32_t* gate) {
    if (gate && *gate == 0) return;
    uint32_t beg = tid * fan;
    if (beg >= count) return;
    uint32_t end = beg + fan < count ? beg + fan : count;
    uint32_t out_key = NO_KEY;
    uint32_t key = NO_KEY;
    bool head = true;
    XYZZ<F> acc;
    set_inf(acc);
    // One add site per element keeps the warp converged: a run that starts here begins from the
    // bucket's current sum and is stored back when the run ends (no add at the flush).
    for (uint32_t t = beg; t < end; t++) {
      uint32_t k = keys_in[t];
      if (k != key) {
        if (key != NO_KEY) {
          if (head) bucket_sums[key] = acc;
          else { parts_out[tid] = acc; out_key = key; }
        }
        key = k;
        head = !(t == beg && beg > 0 && keys_in[beg - 1] == k);
        if (k != NO_KEY && head) acc = bucket_sums[k]; else set_inf(acc);
      }
      if (k != NO_KEY) { XYZZ<F> q = parts_in[t]; xyzz_add_ilp(acc, q); }
    }
    if (key != NO_KEY) {
      if (head) bucket_sums[key] = acc;
      else { parts_out[tid] = acc; out_key = key; }
    }
    keys_out[tid] = out_key;
  }
};

// The fix-up in ONE launch for the common case (what G2 uses, whose per-bucket kernel is register-bound): chunk t of
// the chunked accumulation covers the sorted pairs [t L, (t + 1) L), so bucket b = [s, e) is met by the chunks
// floor(s / L) .. floor((e - 1) / L); the first of them wrote the head of the sum to bucket_sums[b], each later one
// left a partial sum that continues b.  Thread b adds those partials -- at most `cap / L` of them: BucketSizeCheck
// raises *big beforehand when a bucket is larger than cap (skewed scalars), this launch then does nothing and the
// balanced tree of FixupLevel launches, gated the other way, does the work.
struct BucketSizeCheck {
  static const char* name() { return "bucket_size_check"; }
  static ZK_HD void run(uint32_t tid, uint32_t nb, const uint32_t* offsets, uint32_t cap, uint32_t* big) {
    if (tid < nb && offsets[tid + 1] - offsets[tid] > cap) zk_atomic_or(big, 1u);
  }
};
template <class C> struct FixupDirect {
  typedef typename C::F F;
  static const char* name() { return "fixup_direct"; }
  static ZK_HD void run(uint32_t tid, uint32_t nb, uint32_t L, const uint32_t* offsets, const XYZZ<F>* partials,
                        XYZZ<F>* bucket_sums, const uint32_t* big) {
    if (tid >= nb || *big) return;
    const uint32_t s = offsets[tid], e = offsets[tid + 1];
    if (s == e) return;
    const uint32_t t0 = s / L, t1 = (e - 1) / L;
    if (t1 == t0) return;
    XYZZ<F> acc = bucket_sums[tid];
    for (uint32_t t = t0 + 1; t <= t1; t++) { XYZZ<F> q = partials[t]; xyzz_add_ilp(acc, q); }
    bucket_sums[tid] = acc;
  }
};

// ---------------------------------------------------------------- batched-affine pre-reduction
// (on by default for large G1 sets with precomputed slabs, see msm_default_batch_rounds)
// Round r halves every bucket: neighbours (2j, 2j+1) of a bucket are added in AFFINE coordinates,
//   lambda = (y2 - y1) / (x2 - x1),  x3 = lambda^2 - x1 - x2,  y3 = lambda (x1 - x3) - y1      (macros.rs:109-152)
// which is what the reference does per addition -- but the inversions of T independent additions are shared by
// Montgomery's trick: prefix products of the denominators on the way forward, ONE inversion, and two
// multiplications per addition on the way back.  6 field multiplications per addition instead of the 10 of an
// XYZZ mixed add; the single inversion (division steps in batches of 30, fp.cuh) costs ~5 % of a 128-addition batch.
// Every exceptional case keeps the reference semantics: AtInfinity operands and odd leftovers are copied,
// P + P takes the tangent (denominator 2y, macros.rs:57-108), P + (-P) gives AtInfinity (macros.rs:53-56);
// those use the denominator 1 so that the shared product never vanishes.
struct PairCount {   // cnt[b] = ceil(size_b / 2)
  static const char* name() { return "pair_count"; }
  static ZK_HD void run(uint32_t tid, uint32_t nb, const uint32_t* off_in, uint32_t* cnt) {
    if (tid >= nb) return;
    cnt[tid] = (off_in[tid + 1] - off_in[tid] + 1) / 2;
  }
};

template <class C, bool FIRST> struct BatchedAddRound {
  typedef typename C::F F;
  static const char* name() { return FIRST ? "batched_add_first" : "batched_add"; }

  // operands of output o: items i1 = off_in[b] + 2j and i1 + 1 (if it exists), b = bucket of o, j = o - off_out[b]
  struct Where { uint32_t i1; bool has2; };
  static ZK_HD Where where(uint32_t o, uint32_t b, const uint32_t* off_in, const uint32_t* off_out) {
    uint32_t i1 = off_in[b] + 2 * (o - off_out[b]);
    return Where{i1, i1 + 1 < off_in[b + 1]};
  }
  // kind 0 add, 1 double, 2 copy p1, 3 copy p2, 4 infinity; d = the denominator (1 for the kinds without one).
  // The common case -- two points with different non-zero x -- is decided from x alone (d = x2 - x1 is already set).
  static ZK_HD bool common_case(bool has2, const F& x1, const F& x2, const F& d) {
    return has2 && !fis_zero(d) && !fis_zero(x1) && !fis_zero(x2);
  }
  static ZK_HD int classify_rare(F& d, bool has2, const F& x1, const F& y1, const F& x2, const F& y2) {
    const bool inf1 = fis_zero(x1) && fis_zero(y1), inf2 = fis_zero(x2) && fis_zero(y2);
    int kind;
    if (!has2 || inf2) kind = 2;
    else if (inf1) kind = 3;
    else if (!fis_zero(d)) return 0;                               // x differs (one x is zero: the points (0, +-2))
    else if (feq(y1, y2) && !fis_zero(y1)) { fdbl(d, y1); return 1; }
    else kind = 4;
    fset_one(d);
    return kind;
  }

  // thread tid owns outputs [tid T, (tid + 1) T); its prefix products sit at stride kBlockT so that the 32 lanes
  // of a warp read and write neighbouring elements
  static constexpr uint32_t kBlockT = 128;
  static ZK_HD size_t prefix_at(uint32_t tid, uint32_t T, uint32_t i) {
    return (size_t)(tid / kBlockT) * kBlockT * T + (size_t)i * kBlockT + tid % kBlockT;
  }

  // Locating one output's operands takes two dependent steps in the first round (sorted entry -> point address),
  // so both passes run a three-stage software pipeline: entries of output o + 2 (Loc), x coordinates of o + 1
  // (Item + registers), arithmetic of o -- in-order issue would otherwise stall every warp on the entry load.
  struct Loc { uint32_t b, i1; bool has2; Entry e1, e2; };
  struct Item {
    uint32_t b;      // bucket
    bool has2, neg1, neg2;
    const Affine<F>* p1;
    const Affine<F>* p2;
  };
  static ZK_HD void find(Loc& l, uint32_t o, uint32_t b, const uint32_t* off_in, const uint32_t* off_out, const Entry* entries) {
    Where w = where(o, b, off_in, off_out);
    l.b = b; l.i1 = w.i1; l.has2 = w.has2;
    if (FIRST) { l.e1 = entries[w.i1]; l.e2 = entries[w.has2 ? w.i1 + 1 : w.i1]; }
  }
  static ZK_HD void resolve(Item& it, const Loc& l, const Affine<F>* src) {
    it.b = l.b; it.has2 = l.has2;
    if (FIRST) {
      it.p1 = src + (l.e1.val & 0x7fffffffu); it.neg1 = (l.e1.val >> 31) != 0;
      it.p2 = src + (l.e2.val & 0x7fffffffu); it.neg2 = (l.e2.val >> 31) != 0;
    } else {
      it.p1 = src + l.i1; it.p2 = src + (l.has2 ? l.i1 + 1 : l.i1);
      it.neg1 = it.neg2 = false;
    }
  }

  // Forward pass: d_o = denominator of output o, prefix_o = d_beg ... d_(o-1); both are stored (dstride apart) so
  // that the way back starts each output from two coalesced loads and fetches the four coordinates while the
  // first two multiplications run.
  // In the FIRST round -- bound by its random gathers, not by the multiplier -- the denominators are NOT kept:
  // d = x2 - x1 is recomputed from the coordinates the way back loads anyway (one subtraction instead of 48 B
  // written and read back per addition) and the running inverse is advanced at the end of an iteration, once the
  // coordinates have arrived.  Measured at 2^20: first round 2.70 -> 2.63 ms; the streamed rounds lose 3.5 % with
  // the same change (their loads are coalesced and cheap), so they keep the stored denominators.
#if defined(ZK_BATCH_NO_DSTORE)
  static constexpr bool kNoDStore = true;
#elif defined(ZK_BATCH_DSTORE)
  static constexpr bool kNoDStore = false;
#else
  static constexpr bool kNoDStore = FIRST;
#endif
#if defined(__CUDA_ARCH__) && defined(ZK_BATCH_RING)
  // ---- first round with the operands staged through shared memory (cp.async ring), G1 only -------------------
  // The first round gathers its operands from the precomputed tables at random; with the operands of ONE output
  // ahead in registers the warps sit on the long scoreboard (27 % of the stall samples, profiles/).  Here every
  // thread keeps a private ring in shared memory that cp.async fills 4 outputs ahead on the way forward (x1, x2: 6
  // 16-byte chunks per output) and 2 outputs ahead on the way back (x1, y1, x2, y2: 12 chunks), no registers held.
  // Layout: chunk slot cs of thread t at ((cs * 128 + t) * 16) bytes: 24 chunk slots = 48 KB per block.
  static __device__ __forceinline__ void cp16(uint32_t smem_addr, const void* g) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_addr), "l"(g));
  }
  static __device__ __forceinline__ void cp48(uint32_t smem_base, uint32_t cs, const void* g) {
    const char* c = (const char*)g;
    cp16(smem_base + (cs * 128u) * 16u, c);
    cp16(smem_base + ((cs + 1) * 128u) * 16u, c + 16);
    cp16(smem_base + ((cs + 2) * 128u) * 16u, c + 32);
  }
  static __device__ __forceinline__ void ld48(F& r, const uint4* ring, uint32_t cs) {
    uint4 a = ring[cs * 128u + threadIdx.x], b = ring[(cs + 1) * 128u + threadIdx.x], c = ring[(cs + 2) * 128u + threadIdx.x];
    r.v[0] = a.x; r.v[1] = a.y; r.v[2] = a.z; r.v[3] = a.w;
    r.v[4] = b.x; r.v[5] = b.y; r.v[6] = b.z; r.v[7] = b.w;
    r.v[8] = c.x; r.v[9] = c.y; r.v[10] = c.z; r.v[11] = c.w;
  }
  static __device__ void run_ring(uint32_t tid, MsmPlan p, const uint32_t* off_in, const uint32_t* off_out, const Entry* entries,
                                  const Affine<F>* src, Affine<F>* dst, F* prefix) {
    __shared__ uint4 ring[24 * 128];
    const uint32_t sbase = (uint32_t)__cvta_generic_to_shared(ring) + threadIdx.x * 16u;
    const uint32_t total = off_out[p.nb], T = p.batch_T;
    const uint64_t beg64 = (uint64_t)tid * T;
    if (beg64 >= total) return;
    const uint32_t beg = (uint32_t)beg64, end = beg + T < total ? beg + T : total;
    uint32_t lo = 0, hi = p.nb;
    while (hi - lo > 1) { uint32_t mid = (lo + hi) / 2; if (off_out[mid] <= beg) lo = mid; else hi = mid; }
    const uint32_t b_first = lo;
    constexpr uint32_t DF = 4, DB = 2;
    // meta bits of the outputs in flight: slot j holds has2 | neg1 << 1 | neg2 << 2 at bits [4 j, 4 j + 3)
    uint32_t meta = 0;
    auto set_meta = [&](uint32_t slot, const Loc& l) {
      const uint32_t m = (l.has2 ? 1u : 0u) | ((l.e1.val >> 31) << 1) | ((l.e2.val >> 31) << 2);
      meta = (meta & ~(7u << (4 * slot))) | (m << (4 * slot));
    };
    // ---------------- forward
    uint32_t bi = b_first;        // bucket cursor of the issue stage
    Loc loc;
    for (uint32_t j = 0; j < DF; j++) {
      if (beg + j < end) {
        while (beg + j >= off_out[bi + 1]) bi++;
        find(loc, beg + j, bi, off_in, off_out, entries);
        cp48(sbase, j * 6, &src[loc.e1.val & 0x7fffffffu].x);
        cp48(sbase, j * 6 + 3, &src[loc.e2.val & 0x7fffffffu].x);
        set_meta(j, loc);
      }
      asm volatile("cp.async.commit_group;");
    }
    if (beg + DF < end) {
      while (beg + DF >= off_out[bi + 1]) bi++;
      find(loc, beg + DF, bi, off_in, off_out, entries);
    }
    F prod;
    fset_one(prod);
    uint32_t bc = b_first;        // bucket cursor of the consume stage (rare cases re-read their entries)
    uint64_t rare_lo = 0, rare_hi = 0;
    for (uint32_t o = beg; o < end; o++) {
      const uint32_t slot = (o - beg) % DF;
      asm volatile("cp.async.wait_group %0;" ::"n"(DF - 1));
      F x1, x2;
      ld48(x1, ring, slot * 6);
      ld48(x2, ring, slot * 6 + 3);
      const uint32_t m = (meta >> (4 * slot)) & 7u;
      if (o + DF < end) {
        cp48(sbase, slot * 6, &src[loc.e1.val & 0x7fffffffu].x);
        cp48(sbase, slot * 6 + 3, &src[loc.e2.val & 0x7fffffffu].x);
        set_meta(slot, loc);
        if (o + DF + 1 < end) {
          while (o + DF + 1 >= off_out[bi + 1]) bi++;
          find(loc, o + DF + 1, bi, off_in, off_out, entries);
        }
      }
      asm volatile("cp.async.commit_group;");
      const bool has2 = (m & 1u) != 0;
      F d;
      fsub(d, x2, x1);
      const bool rare = !common_case(has2, x1, x2, d);
      rare_hi = (rare_hi << 1) | (rare_lo >> 63);
      rare_lo = (rare_lo << 1) | (rare ? 1u : 0u);
      if (rare) {
        while (o >= off_out[bc + 1]) bc++;
        Loc lr;
        find(lr, o, bc, off_in, off_out, entries);
        F y1 = src[lr.e1.val & 0x7fffffffu].y, y2 = src[lr.e2.val & 0x7fffffffu].y;
        fcneg(y1, y1, (m & 2u) != 0);
        fcneg(y2, y2, (m & 4u) != 0);
        classify_rare(d, has2, x1, y1, x2, y2);
      }
      prefix[prefix_at(tid, T, o - beg)] = prod;
      fmul(prod, prod, d);
    }
    asm volatile("cp.async.wait_group 0;");
    F inv;
    finv(inv, prod);
    // ---------------- backward: outputs end - 1 .. beg; slot of output o is (end - 1 - o) % DB, 12 chunks each
    while (end - 1 >= off_out[bi + 1]) bi++;          // bucket of the last output
    while (end - 1 < off_out[bi]) bi--;
    for (uint32_t j = 0; j < DB; j++) {
      if (end >= beg + 1 + j) {
        const uint32_t o = end - 1 - j;
        while (o < off_out[bi]) bi--;
        find(loc, o, bi, off_in, off_out, entries);
        const Affine<F>* p1 = &src[loc.e1.val & 0x7fffffffu];
        const Affine<F>* p2 = &src[loc.e2.val & 0x7fffffffu];
        cp48(sbase, j * 12, &p1->x); cp48(sbase, j * 12 + 3, &p1->y);
        cp48(sbase, j * 12 + 6, &p2->x); cp48(sbase, j * 12 + 9, &p2->y);
        set_meta(j, loc);
      }
      asm volatile("cp.async.commit_group;");
    }
    if (end >= beg + 1 + DB) {
      const uint32_t o = end - 1 - DB;
      while (o < off_out[bi]) bi--;
      find(loc, o, bi, off_in, off_out, entries);
    }
    F pk = prefix[prefix_at(tid, T, end - 1 - beg)];
    for (uint32_t o = end; o-- > beg;) {
      const uint32_t slot = (end - 1 - o) % DB;
      asm volatile("cp.async.wait_group %0;" ::"n"(DB - 1));
      F x1, y1, x2, y2, dinv, lam, t, d;
      ld48(x1, ring, slot * 12); ld48(y1, ring, slot * 12 + 3);
      ld48(x2, ring, slot * 12 + 6); ld48(y2, ring, slot * 12 + 9);
      const uint32_t m = (meta >> (4 * slot)) & 7u;
      if (o >= beg + DB) {                              // output o - DB takes this slot
        const Affine<F>* p1 = &src[loc.e1.val & 0x7fffffffu];
        const Affine<F>* p2 = &src[loc.e2.val & 0x7fffffffu];
        cp48(sbase, slot * 12, &p1->x); cp48(sbase, slot * 12 + 3, &p1->y);
        cp48(sbase, slot * 12 + 6, &p2->x); cp48(sbase, slot * 12 + 9, &p2->y);
        set_meta(slot, loc);
        if (o >= beg + DB + 1) {
          const uint32_t on = o - DB - 1;
          while (on < off_out[bi]) bi--;
          find(loc, on, bi, off_in, off_out, entries);
        }
      }
      asm volatile("cp.async.commit_group;");
      fmul(dinv, inv, pk);
      if (o > beg) pk = prefix[prefix_at(tid, T, o - 1 - beg)];
      const bool has2 = (m & 1u) != 0;
      fcneg(y1, y1, (m & 2u) != 0);
      fcneg(y2, y2, (m & 4u) != 0);
      int kind = 0;
      const bool rare = (rare_lo & 1u) != 0;
      rare_lo = (rare_lo >> 1) | (rare_hi << 63);
      rare_hi >>= 1;
      fsub(d, x2, x1);
      if (rare) kind = classify_rare(d, has2, x1, y1, x2, y2);
      fsub(t, y2, y1);
      if (kind == 1) { fmul(t, x1, x1); fdbl(lam, t); fadd(t, lam, t); x2 = x1; }
      fmul(lam, t, dinv);
      fmul(inv, inv, d);
      Affine<F> r;
      fmul(t, lam, lam);
      fsub(t, t, x1);
      fsub(r.x, t, x2);
      fsub(t, x1, r.x);
      fmul(t, lam, t);
      fsub(r.y, t, y1);
      if (kind >= 2) {
        if (kind == 2) { r.x = x1; r.y = y1; }
        else if (kind == 3) { r.x = x2; r.y = y2; }
        else set_inf(r);
      }
      dst[o] = r;
    }
    asm volatile("cp.async.wait_group 0;");
  }
#endif

  static ZK_HD void run(uint32_t tid, MsmPlan p, const uint32_t* off_in, const uint32_t* off_out, const Entry* entries,
                        const Affine<F>* src, Affine<F>* dst, F* prefix, size_t dstride, Entry* entries_out) {
#if defined(__CUDA_ARCH__) && defined(ZK_BATCH_RING)
    if (FIRST && sizeof(F) == 48 && entries_out == nullptr) { run_ring(tid, p, off_in, off_out, entries, src, dst, prefix); return; }
#endif
    const uint32_t total = off_out[p.nb], T = p.batch_T;
    const uint64_t beg64 = (uint64_t)tid * T;
    if (beg64 >= total) return;
    const uint32_t beg = (uint32_t)beg64, end = beg + T < total ? beg + T : total;
    uint32_t lo = 0, hi = p.nb;                       // largest b with off_out[b] <= beg
    while (hi - lo > 1) { uint32_t mid = (lo + hi) / 2; if (off_out[mid] <= beg) lo = mid; else hi = mid; }
    uint32_t b = lo;                                  // bucket of the output the Loc stage is at
    F prod;
    fset_one(prod);
    Loc loc;
    Item cur, nxt;
    find(loc, beg, b, off_in, off_out, entries);
    resolve(cur, loc, src);
    F x1 = cur.p1->x, x2 = cur.p2->x;
    if (beg + 1 < end) {
      while (beg + 1 >= off_out[b + 1]) b++;
      find(loc, beg + 1, b, off_in, off_out, entries);
    }
    uint64_t rare_lo = 0, rare_hi = 0;   // stack of "not the common case" flags, one bit per output (T <= 128)
    for (uint32_t o = beg; o < end; o++) {
      F nx1, nx2;
      if (o + 1 < end) {
        resolve(nxt, loc, src);
        nx1 = nxt.p1->x; nx2 = nxt.p2->x;
        if (o + 2 < end) {
          while (o + 2 >= off_out[b + 1]) b++;
          find(loc, o + 2, b, off_in, off_out, entries);
        }
      }
      F d;
      fsub(d, x2, x1);
      const bool rare = !common_case(cur.has2, x1, x2, d);
      rare_hi = (rare_hi << 1) | (rare_lo >> 63);
      rare_lo = (rare_lo << 1) | (rare ? 1u : 0u);
      if (rare) {
        F y1 = cur.p1->y, y2 = cur.p2->y;
        fcneg(y1, y1, cur.neg1);
        fcneg(y2, y2, cur.neg2);
        classify_rare(d, cur.has2, x1, y1, x2, y2);
      }
      const size_t at = prefix_at(tid, T, o - beg);
      prefix[at] = prod;
      if (!kNoDStore) prefix[dstride + at] = d;
      fmul(prod, prod, d);
      if (o + 1 < end) { cur = nxt; x1 = nx1; x2 = nx2; }
    }
    F inv;
    finv(inv, prod);
    // backward: inverse of each denominator, then the affine formulas.  cur is the last output's item, b its bucket.
    if (kNoDStore) {
      size_t at = prefix_at(tid, T, end - 1 - beg);
      F pk = prefix[at];
      if (end - 1 > beg) {
        while (end - 2 < off_out[b]) b--;
        find(loc, end - 2, b, off_in, off_out, entries);
      }
      for (uint32_t o = end; o-- > beg;) {
        const Item me = cur;
        F y1 = me.p1->y, y2 = me.p2->y, dinv, lam, t, d;
        x1 = me.p1->x; x2 = me.p2->x;
        fmul(dinv, inv, pk);
        if (FIRST) {
          fcneg(y1, y1, me.neg1);
          fcneg(y2, y2, me.neg2);
        }
        int kind = 0;
        const bool rare = (rare_lo & 1u) != 0;
        rare_lo = (rare_lo >> 1) | (rare_hi << 63);
        rare_hi >>= 1;
        fsub(d, x2, x1);
        if (rare) kind = classify_rare(d, me.has2, x1, y1, x2, y2);   // d becomes 1 or 2 y, as in the forward pass
        if (o > beg) {   // next output: its prefix product ahead of its use
          resolve(cur, loc, src);
          at = prefix_at(tid, T, o - 1 - beg);
          pk = prefix[at];
          if (o - 1 > beg) {
            while (o - 2 < off_out[b]) b--;
            find(loc, o - 2, b, off_in, off_out, entries);
          }
        }
        fsub(t, y2, y1);
        if (kind == 1) { fmul(t, x1, x1); fdbl(lam, t); fadd(t, lam, t); x2 = x1; }   // tangent: 3 x^2 / 2 y
        fmul(lam, t, dinv);
        fmul(inv, inv, d);
        Affine<F> r;
        fmul(t, lam, lam);
        fsub(t, t, x1);
        fsub(r.x, t, x2);
        fsub(t, x1, r.x);
        fmul(t, lam, t);
        fsub(r.y, t, y1);
        if (kind >= 2) {
          if (kind == 2) { r.x = x1; r.y = y1; }
          else if (kind == 3) { r.x = x2; r.y = y2; }
          else set_inf(r);
        }
        dst[o] = r;
        if (entries_out) { Entry e; e.key = me.b; e.val = o; entries_out[o] = e; }
      }
    } else {
      size_t at = prefix_at(tid, T, end - 1 - beg);
      F pk = prefix[at], d = prefix[dstride + at];
      if (end - 1 > beg) {
        while (end - 2 < off_out[b]) b--;
        find(loc, end - 2, b, off_in, off_out, entries);
      }
      for (uint32_t o = end; o-- > beg;) {
        const Item me = cur;
        F y1 = me.p1->y, y2 = me.p2->y, dinv, lam, t;
        x1 = me.p1->x; x2 = me.p2->x;
        fmul(dinv, inv, pk);
        fmul(inv, inv, d);
        if (FIRST) {
          fcneg(y1, y1, me.neg1);
          fcneg(y2, y2, me.neg2);
        }
        int kind = 0;
        const bool rare = (rare_lo & 1u) != 0;
        rare_lo = (rare_lo >> 1) | (rare_hi << 63);
        rare_hi >>= 1;
        if (rare) {
          fsub(t, x2, x1);   // the raw difference decides the case (the stored d is 1 or 2y here)
          kind = classify_rare(t, me.has2, x1, y1, x2, y2);
        }
        if (o > beg) {   // next output: its prefix / denominator three multiplications ahead of their use
          resolve(cur, loc, src);
          at = prefix_at(tid, T, o - 1 - beg);
          pk = prefix[at]; d = prefix[dstride + at];
          if (o - 1 > beg) {
            while (o - 2 < off_out[b]) b--;
            find(loc, o - 2, b, off_in, off_out, entries);
          }
        }
        fsub(t, y2, y1);
        if (kind == 1) { fmul(t, x1, x1); fdbl(lam, t); fadd(t, lam, t); x2 = x1; }   // tangent: 3 x^2 / 2 y
        fmul(lam, t, dinv);
        Affine<F> r;
        fmul(t, lam, lam);
        fsub(t, t, x1);
        fsub(r.x, t, x2);
        fsub(t, x1, r.x);
        fmul(t, lam, t);
        fsub(r.y, t, y1);
        if (kind >= 2) {
          if (kind == 2) { r.x = x1; r.y = y1; }
          else if (kind == 3) { r.x = x2; r.y = y2; }
          else set_inf(r);
        }
        dst[o] = r;
        if (entries_out) { Entry e; e.key = me.b; e.val = o; entries_out[o] = e; }
      }
    }
  }
};

// p = s * p for a small scalar (left-to-right double-and-add)
template <class F> ZK_HD void xyzz_mul_small(XYZZ<F>& p, uint32_t s) {
  if (s == 0) { set_inf(p); return; }
  XYZZ<F> base = p;
  int top = 31;
  while (!((s >> top) & 1)) top--;
  for (int b = top - 1; b >= 0; b--) {
    xyzz_dbl_ilp(p);
    if ((s >> b) & 1) xyzz_add_ilp(p, base);
  }
}

// thread (win, k): out = sum_{i<K} (g + i + 1) * bucket[win*B + k*K + i], g = global index of local bucket k K (a chain
// never straddles a stripe: K divides the stripe length); empty buckets were never written
template <class C> struct BucketReduce {
  typedef typename C::F F;
  static const char* name() { return "bucket_reduce"; }
  static ZK_HD void run(uint32_t tid, MsmPlan p, const uint32_t* offsets, const XYZZ<F>* bucket_sums, XYZZ<F>* out) {
    uint32_t chunks = p.B / p.K;
    if (tid >= p.nwin * chunks) return;
    uint32_t win = tid / chunks, k = tid % chunks;
    size_t base = (size_t)win * p.B + (size_t)k * p.K;
    XYZZ<F> run, acc;
    set_inf(run);
    set_inf(acc);
    for (int i = (int)p.K - 1; i >= 0; i--) {
      if (offsets[base + i] != offsets[base + i + 1]) { XYZZ<F> q = bucket_sums[base + i]; xyzz_add_ilp(run, q); }
      xyzz_add_ilp(acc, run);
    }
    xyzz_mul_small(run, msm_global_bucket(p, k * p.K));
    xyzz_add_ilp(acc, run);
    out[tid] = acc;
  }
};

// level of the pairwise tree: arr is nwin rows of `m` points (row pitch `pitch`); row[i] += row[i + half]
template <class C> struct PairSum {
  typedef typename C::F F;
  static const char* name() { return "pair_sum"; }
  static ZK_HD void run(uint32_t tid, uint32_t nwin, uint32_t pitch, uint32_t m, uint32_t half, XYZZ<F>* arr) {
    if (tid >= nwin * half) return;
    uint32_t win = tid / half, i = tid % half;
    if (i + half >= m) return;
    XYZZ<F>* row = arr + (size_t)win * pitch;
    XYZZ<F> a = row[i], b = row[i + half];
    xyzz_add_ilp(a, b);
    row[i] = a;
  }
};

template <class F> ZK_HD void store_canonical(uint32_t* out, const Affine<F>& a);
template <> ZK_HD void store_canonical<Fp>(uint32_t* out, const Affine<Fp>& a) {
  ffrom_mont(out, a.x); ffrom_mont(out + 12, a.y);
}
template <> ZK_HD void store_canonical<Fp2>(uint32_t* out, const Affine<Fp2>& a) {
  ffrom_mont(out, a.x.c0); ffrom_mont(out + 12, a.x.c1); ffrom_mont(out + 24, a.y.c0); ffrom_mont(out + 36, a.y.c1);
}
template <class F> ZK_HD void load_canonical(Affine<F>& a, const uint32_t* in);
template <> ZK_HD void load_canonical<Fp>(Affine<Fp>& a, const uint32_t* in) {
  fto_mont(a.x, in); fto_mont(a.y, in + 12);
}
template <> ZK_HD void load_canonical<Fp2>(Affine<Fp2>& a, const uint32_t* in) {
  fto_mont(a.x.c0, in); fto_mont(a.x.c1, in + 12); fto_mont(a.y.c0, in + 24); fto_mont(a.y.c1, in + 36);
}

// Horner over the window sums (row w of arr, element 0), then optional canonical affine output.
// out_xyzz (Montgomery XYZZ, for multi-GPU partials) and out_affine (canonical limbs + flag) may be null.
template <class C> struct Finish {
  typedef typename C::F F;
  static const char* name() { return "finish"; }
  static ZK_HD void run(uint32_t tid, uint32_t nwin, uint32_t pitch, uint32_t c, const XYZZ<F>* arr,
                        XYZZ<F>* out_xyzz, uint32_t* out_affine, uint32_t* out_inf) {
    if (tid != 0) return;
    XYZZ<F> acc = arr[(size_t)(nwin - 1) * pitch];
    for (int w = (int)nwin - 2; w >= 0; w--) {
      for (uint32_t i = 0; i < c; i++) xyzz_dbl_ilp(acc);
      XYZZ<F> q = arr[(size_t)w * pitch];
      xyzz_add_ilp(acc, q);
    }
    if (out_xyzz) *out_xyzz = acc;
    if (out_affine) {
      Affine<F> a;
      xyzz_to_affine(a, acc);
      store_canonical<F>(out_affine, a);
      *out_inf = is_inf(acc) ? 1u : 0u;
    }
  }
};

// A partial whose MSM saw an out-of-range scalar is exported POISONED (ZZ = 0, ZZZ = all ones: not a field
// element), so that the error travels with the blob through any gather and the combine step reports it.
template <class F> ZK_HD void poison_partial(XYZZ<F>& q) {
  uint32_t* w = reinterpret_cast<uint32_t*>(&q);
  constexpr int NL = sizeof(F) / 4;
  for (int i = 0; i < NL; i++) { w[2 * NL + i] = 0; w[3 * NL + i] = 0xffffffffu; }
}
template <class F> ZK_HD bool is_poisoned(const XYZZ<F>& q) {
  const uint32_t* w = reinterpret_cast<const uint32_t*>(&q);
  constexpr int NL = sizeof(F) / 4;
  uint32_t all = 0xffffffffu, zz = 0;
  for (int i = 0; i < NL; i++) { all &= w[3 * NL + i]; zz |= w[2 * NL + i]; }
  return all == 0xffffffffu && zz == 0;
}
template <class C> struct PoisonPartial {   // after Finish, when the partial leaves the context without a status read
  typedef typename C::F F;
  static const char* name() { return "poison_partial"; }
  static ZK_HD void run(uint32_t tid, const uint32_t* err, XYZZ<F>* out) {
    if (tid == 0 && *err) poison_partial(*out);
  }
};

// sum of k XYZZ partial results (multi-GPU combine) -> canonical affine; *err |= ERR_SCALAR_RANGE for a poisoned partial
template <class C> struct CombinePartials {
  typedef typename C::F F;
  static const char* name() { return "combine_partials"; }
  // stride_words: distance between consecutive partials in 32-bit words (sizeof(XYZZ<F>) / 4 for a packed array)
  static ZK_HD void run(uint32_t tid, uint32_t k, const XYZZ<F>* parts, uint32_t stride_words, uint32_t* out_affine, uint32_t* out_inf,
                        uint32_t* err) {
    if (tid != 0) return;
    XYZZ<F> acc;
    set_inf(acc);
    for (uint32_t i = 0; i < k; i++) {
      XYZZ<F> q = *reinterpret_cast<const XYZZ<F>*>(reinterpret_cast<const uint32_t*>(parts) + (size_t)i * stride_words);
      if (is_poisoned(q)) { zk_atomic_or(err, ERR_SCALAR_RANGE); continue; }
      xyzz_add_ilp(acc, q);
    }
    Affine<F> a;
    xyzz_to_affine(a, acc);
    store_canonical<F>(out_affine, a);
    *out_inf = is_inf(acc) ? 1u : 0u;
  }
};

// ---------------------------------------------------------------- point-set preparation
// canonical affine limbs (+ optional infinity flags) -> Montgomery affine, slab 0
template <class C> struct LoadPoints {
  typedef typename C::F F;
  static const char* name() { return "load_points"; }
  static ZK_HD void run(uint32_t tid, uint32_t n, const uint32_t* canon, const uint8_t* inf, Affine<F>* out) {
    if (tid >= n) return;
    Affine<F> a;
    if (inf && inf[tid]) set_inf(a);
    else load_canonical<F>(a, canon + (size_t)tid * C::AFF_LIMBS);
    out[tid] = a;
  }
};

// slab w = 2^width(w-1) * slab (w-1), w = 1 .. W-1, i.e. 2^pos(w) P (one thread per point walks all levels); the first
// `wide` windows are c bits wide, the others c - 1 (Digits)
template <class C> struct PrecomputeSlabs {
  typedef typename C::F F;
  static const char* name() { return "precompute_slabs"; }
  static ZK_HD void run(uint32_t tid, uint32_t n, uint32_t stride, uint32_t c, uint32_t W, uint32_t wide, Affine<F>* pts) {
    if (tid >= n) return;
    Affine<F> a = pts[tid];
    for (uint32_t w = 1; w < W; w++) {
      const uint32_t t = w - 1 < wide ? c : c - 1;
      if (!is_inf(a)) {
        XYZZ<F> x;
        xyzz_mdbl(x, a);
        for (uint32_t i = 1; i < t; i++) xyzz_dbl(x);
        xyzz_to_affine(a, x);
      }
      pts[(size_t)w * stride + tid] = a;
    }
  }
};

// ---------------------------------------------------------------- subgroup membership (opt-in at load time)
// ZKMSM_SUBGROUP sets fold a scalar s > (r-1)/2 to (r - s)(-P), which is s P only when P has order r; both curves
// have cofactors, so a point that is on the curve but outside the subgroup would silently give another result than
// the reference's raw multiple (macros.rs:10-21).  With ZKMSM_CHECK_SUBGROUP the load verifies r P = AtInfinity for
// every point (MSB-first double-and-add over the 255 bits of r, mixed additions) and counts the failures.
template <class C> struct SubgroupCheck {
  typedef typename C::F F;
  static const char* name() { return "subgroup_check"; }
  static ZK_HD void run(uint32_t tid, uint32_t n, const Affine<F>* pts, uint32_t* bad) {
    if (tid >= n) return;
    const Affine<F> p = pts[tid];
    if (is_inf(p)) return;
    XYZZ<F> acc;
    from_affine(acc, p);
    for (int bit = 253; bit >= 0; bit--) {          // bit 254 is r's top bit: acc starts at P
      xyzz_dbl(acc);
      if ((FR_P[bit >> 5] >> (bit & 31)) & 1u) xyzz_madd(acc, p);
    }
    if (!is_inf(acc)) zk_atomic_add(bad, 1u);
  }
};

// ---------------------------------------------------------------- fixed-base scalar multiplication
// (the reference's `g * k`, impl_scalar_mul_point!, macros.rs:2-32, for many k at once; used for
//  CRS-style point generation, crs.rs:88-116)
// table[j] = 2^j * base, j < 256, built by one thread then converted to affine in parallel
template <class C> struct BaseTableChain {
  typedef typename C::F F;
  static const char* name() { return "base_table_chain"; }
  static ZK_HD void run(uint32_t tid, const uint32_t* base_canon, XYZZ<F>* chain) {
    if (tid != 0) return;
    Affine<F> a;
    load_canonical<F>(a, base_canon);
    XYZZ<F> x;
    from_affine(x, a);
    for (int j = 0; j < 256; j++) { chain[j] = x; xyzz_dbl(x); }
  }
};
template <class C> struct BaseTableAffine {
  typedef typename C::F F;
  static const char* name() { return "base_table_affine"; }
  static ZK_HD void run(uint32_t tid, const XYZZ<F>* chain, Affine<F>* table) {
    if (tid >= 256) return;
    Affine<F> a;
    xyzz_to_affine(a, chain[tid]);
    table[tid] = a;
  }
};
// out[i] = scalars[i] * base as Montgomery affine ((0,0) for infinity); full 256-bit scalars
template <class C> struct FixedBaseMul {
  typedef typename C::F F;
  static const char* name() { return "fixed_base_mul"; }
  static ZK_HD void run(uint32_t tid, uint32_t n, const uint32_t* scalars, const Affine<F>* table, Affine<F>* out) {
    if (tid >= n) return;
    XYZZ<F> acc;
    set_inf(acc);
    for (int limb = 0; limb < SCALAR_LIMBS; limb++) {
      uint32_t w = scalars[(size_t)tid * SCALAR_LIMBS + limb];
      for (int b = 0; b < 32; b++) {
        if ((w >> b) & 1) { Affine<F> q = table[limb * 32 + b]; xyzz_madd(acc, q); }
      }
    }
    Affine<F> a;
    xyzz_to_affine(a, acc);
    out[tid] = a;
  }
};
// Montgomery affine -> canonical limbs + infinity flags (inverse of LoadPoints)
template <class C> struct StorePoints {
  typedef typename C::F F;
  static const char* name() { return "store_points"; }
  static ZK_HD void run(uint32_t tid, uint32_t n, const Affine<F>* in, uint32_t* canon, uint8_t* inf) {
    if (tid >= n) return;
    Affine<F> a = in[tid];
    store_canonical<F>(canon + (size_t)tid * C::AFF_LIMBS, a);
    if (inf) inf[tid] = is_inf(a) ? 1 : 0;
  }
};

// ---------------------------------------------------------------- planning
// Tuning / cross-check switches.  The product reads the environment ONCE per context (zkmsm_create) and lets
// zkmsm_set_option override single fields; the CPU emulation reads it per call.  -1 / 0 = automatic.
struct MsmTuning {
  int batch_rounds = -1;   // ZKMSM_BATCH_ROUNDS: force the number of batched-affine rounds (0 = off)
  int batch_T = 0;         // ZKMSM_BATCH_T: additions sharing one inversion (2..128)
  int batch_g2 = 1;        // ZKMSM_BATCH_G2=0: no batched-affine rounds for G2
  int batch_blocks = 0;    // ZKMSM_BATCH_BLOCKS: resident blocks per SM assumed for the batched kernel (A/B builds)
  int L = 0;               // ZKMSM_L: sorted pairs per accumulate thread
  int K = 0;               // ZKMSM_K: buckets per reduction chain (power of two)
  int no_wave_L = 0;       // ZKMSM_NO_WAVE_L
  int no_coop = 0;         // ZKMSM_NO_COOP: per-thread tail kernels
  int ntt_no_fuse = 0;     // ZKMSM_NTT_NO_FUSE
  int quotient_schoolbook = 0;   // ZKMSM_QUOTIENT_SCHOOLBOOK
  int no_graph = 0;        // ZKMSM_NO_GRAPH: launch kernel by kernel instead of replaying a captured CUDA graph
  int no_bucket_acc = 0;   // ZKMSM_NO_BUCKET_ACC: always the chunked accumulation + fix-up tree
  int acc_G = 0;           // ZKMSM_ACC_G: lanes per bucket of AccumulateBuckets (power of two <= 32)
  int coop_max_chains = 0; // ZKMSM_COOP_MAX_CHAINS: most reduction chains the cooperative kernel takes (default 9472)
  int min_left = 0;        // ZKMSM_MIN_LEFT: batched rounds stop at about this many items per bucket (default 6)
  static int env_int(const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; }
  static MsmTuning from_env() {
    MsmTuning t;
    t.batch_rounds = env_int("ZKMSM_BATCH_ROUNDS", -1);
    t.batch_T = env_int("ZKMSM_BATCH_T", 0);
    t.batch_g2 = env_int("ZKMSM_BATCH_G2", 1);
    t.batch_blocks = env_int("ZKMSM_BATCH_BLOCKS", 0);
    t.L = env_int("ZKMSM_L", 0);
    t.K = env_int("ZKMSM_K", 0);
    t.no_wave_L = getenv("ZKMSM_NO_WAVE_L") ? 1 : 0;
    t.no_coop = getenv("ZKMSM_NO_COOP") ? 1 : 0;
    t.ntt_no_fuse = getenv("ZKMSM_NTT_NO_FUSE") ? 1 : 0;
    t.quotient_schoolbook = getenv("ZKMSM_QUOTIENT_SCHOOLBOOK") ? 1 : 0;
    t.no_graph = getenv("ZKMSM_NO_GRAPH") ? 1 : 0;
    t.no_bucket_acc = getenv("ZKMSM_NO_BUCKET_ACC") ? 1 : 0;
    t.acc_G = env_int("ZKMSM_ACC_G", 0);
    t.coop_max_chains = env_int("ZKMSM_COOP_MAX_CHAINS", 0);
    t.min_left = env_int("ZKMSM_MIN_LEFT", 0);
    return t;
  }
};

// windows needed for 255-bit scalars (254-bit after the half-range fold): the last one absorbs the recoding carry
inline uint32_t msm_windows(uint32_t c, bool half = false) { return (half ? 254u : 255u) / c + 1; }

// Wide (c-bit) windows of a set's layout.  A plain set keeps every window at c bits (the top one is whatever the
// scalar leaves: 255 - (W - 1) c bits).  A precomputed set shares ONE bucket set between all windows, and a short top
// window would pour its n digits into 2^(tb-1) of the 2^(c-1) buckets -- 4096 buckets with 4096 extra entries each
// at 2^24 (c = 22, tb = 12): giant buckets for the per-bucket kernels and, under the bucket-range split, for rank 0.
// Its windows are therefore balanced: `wide` of c bits, the rest of c - 1, together exactly the scalar's bits
// (2^20: 14 x 17 + 16, the layout it always had; 2^24: 2 x 22 + 10 x 21).
inline uint32_t msm_wide_windows(uint32_t c, uint32_t W, bool half, bool precomp) {
  if (!precomp) return W;
  const uint32_t bits = half ? 254u : 255u;
  return bits > W * (c - 1) ? bits - W * (c - 1) : 0u;
}

// Window choice.  Cost model in mixed-add units, calibrated on B200 at n = 2^20 (profiles/): expected
// sorted pairs (a full window contributes n, a top window of tb bits n (1 - 2^-tb), a carry-only top
// window n / 2) plus a per-bucket charge for fix-up + reduction (11 per bucket, times W without a shared
// window: ~4 ns per bucket vs 0.37 ns per mixed add), plus a penalty when a narrow top window funnels its
// pairs into a few giant buckets.
inline uint32_t msm_pick_c(uint32_t n, bool precomp, bool half = false) {
  uint32_t best = 8;
  double best_cost = 1e300;
  for (uint32_t c = 3; c <= 22; c++) {
    uint32_t bits = half ? 254u : 255u, full = bits / c, tb = bits - full * c;
    double W = full + 1, B = (double)(1u << (c - 1));
    double entries = (double)n * (full + (tb == 0 ? 0.5 : 1.0 - 1.0 / (double)(1u << (tb > 30 ? 30 : tb))));
    // per-bucket cost in units of one mixed addition: with precomputed slabs the single bucket set is reduced
    // by the latency-bound cooperative tail (measured crossover ~6), otherwise W sets are work-bound (~11)
    double cost = entries + (precomp ? 6.0 : 11.0) * B * (precomp ? 1.0 : W);
    if (2 * tb < c) cost += 0.5 * (double)n;   // carry-only or narrow top window: a few giant buckets
    if (cost < best_cost) { best_cost = cost; best = c; }
  }
  return best;
}

// true when the sorted-pair count of an MSM of n terms at window c fits the 32-bit positions the kernels use
inline bool msm_fits(uint64_t n, uint32_t c, bool half = false) { return n * msm_windows(c, half) < (1ull << 32); }

// acc_slots: accumulate threads the device keeps resident (SMs x 256 at 2 blocks of 128 per SM), 0 = unknown.
// (rank, world): bucket-range split of a precomputed set over `world` devices (a power of two <= 2^(c-1)); the plan
// then owns 2^(c-1) / world buckets and expects 1 / world of the sorted pairs (the buffers stay sized for all of them).
inline MsmPlan msm_plan(uint32_t n, uint32_t c, bool precomp, uint32_t stride, bool half = false, bool coop_tail = false,
                        uint32_t acc_slots = 0, const MsmTuning& tune = MsmTuning(), uint32_t rank = 0, uint32_t world = 1) {
  MsmPlan p;
  p.n = n;
  p.c = c;
  p.half = half ? 1 : 0;
  p.W = msm_windows(c, half);
  p.a = msm_wide_windows(c, p.W, half, precomp);
  p.B = 1u << (c - 1);
  p.rank = 0;
  p.world_log = 0;
  p.stripe_log = 0;
  p.precomp = precomp ? 1 : 0;
  if (precomp && world > 1 && (world & (world - 1)) == 0 && world <= p.B && rank < world) {
    p.B /= world;
    p.rank = rank;
    while ((1u << p.world_log) < world) p.world_log++;
  }
  p.nwin = precomp ? 1 : p.W;
  p.nb = p.nwin * p.B;
  p.stride = stride;
  const uint64_t entries64 = (uint64_t)n * p.W;
  p.max_entries = entries64 < (1ull << 32) ? (uint32_t)entries64 : 0xffffffffu;   // callers reject the latter (msm_fits)
  p.acc_slots = acc_slots;
  p.acc_G = 0;
  p.acc_cap = 0;
  p.L = p.max_entries >= (1u << 23) ? 32 : (p.max_entries >= (1u << 21) ? 16 : 8);
  if (acc_slots && !tune.no_wave_L) {
    // Wave quantisation: every resident slot runs ceil(threads / slots) threads of L mixed additions one after the
    // other, so with few waves (small n) the last, partly filled wave costs a whole one.  Take the L near the
    // default that minimises waves x L (ties: the larger L, fewer partial sums for the fix-up tree).
    uint32_t best = p.L;
    uint64_t best_cost = ~0ull;
    for (uint32_t l = p.L - p.L / 4; l <= p.L + p.L / 2; l++) {
      uint64_t threads = ((uint64_t)p.max_entries + l - 1) / l, waves = (threads + acc_slots - 1) / acc_slots;
      if (waves * l <= best_cost) { best_cost = waves * l; best = l; }
    }
    p.L = best;
  }
  p.K = 2;                                   // per-thread reduction: ~16k threads, short chains, chip busy
  while (p.K < 64 && p.B / p.K > 16384) p.K *= 2;
  if (p.K > p.B) p.K = p.B;
  p.batch_rounds = 0;                        // see msm_default_batch_rounds()
  p.batch_T = 128;
  if (tune.batch_rounds >= 0 && tune.batch_rounds <= 8) p.batch_rounds = (uint32_t)tune.batch_rounds;
  if (tune.batch_T >= 2 && tune.batch_T <= 128) p.batch_T = (uint32_t)tune.batch_T;   // <= 128: one flag bit per output in two words
  p.coop = 0;
  if (coop_tail && !tune.no_coop) {
    // cooperative reduction: 32 chains per 4-warp block.  Up to ~one block per SM (148 x 32 = 4736 chains) every chain
    // runs at single-block latency; give up (per-thread kernel) when that needs K > 64
    // (2^19 buckets at K = 64: 256 blocks, 0.9 ms against 1.8 ms for the per-thread kernel at K = 32)
    uint32_t k = 2;
    while (k < 64 && (uint64_t)p.nwin * (p.B / k) > 4736) k *= 2;
    const uint64_t coop_max = tune.coop_max_chains > 0 ? (uint64_t)tune.coop_max_chains : 2 * 4736;
    if (k <= p.B && (uint64_t)p.nwin * (p.B / k) <= coop_max) { p.K = k; p.coop = 1; }
  }
  if (tune.L >= 1 && tune.L <= 4096) p.L = (uint32_t)tune.L;   // tuning overrides
  if (tune.K >= 1 && (uint32_t)tune.K <= p.B && (tune.K & (tune.K - 1)) == 0) p.K = (uint32_t)tune.K;
  // stripes of the bucket-range split: 32 reduction chains' worth of buckets (one cooperative block), never more
  // than this rank's share
  {
    uint32_t stripe = 32 * p.K;
    while (stripe > p.B) stripe /= 2;
    while ((1u << p.stripe_log) < stripe) p.stripe_log++;
  }
  p.acc_threads = (p.max_entries + p.L - 1) / p.L;
  if (p.acc_threads == 0) p.acc_threads = 1;
  return p;
}

// sorted pairs this plan expects for uniformly random scalars (the buffers hold max_entries, the worst case)
inline uint64_t msm_expected_entries(const MsmPlan& p) {
  const uint64_t full = (uint64_t)1 << (p.c - 1);
  return p.precomp ? (uint64_t)p.max_entries * p.B / full : p.max_entries;
}

// Lanes per bucket and item cap for AccumulateBuckets when about `left` items in total are expected in the plan's
// buckets: as many lanes as still fit ONE wave of the device (a second wave costs as much as the first: measured
// 0.31 ms with 4 lanes against 0.35 ms with 8 for 8192 buckets of 60 items), never more lanes than half the items
// of an average bucket, at most ~256 items per lane before the chunked fallback is the better kernel.
inline void msm_pick_bucket_acc(MsmPlan& p, uint64_t left, const MsmTuning& tune) {
  p.acc_G = 0;
  p.acc_cap = 0;
  if (tune.no_bucket_acc || p.nb == 0) return;
  const uint64_t avg = left / p.nb + 1;
  uint32_t G = 1;
  const uint64_t slots = p.acc_slots ? p.acc_slots : 32768;
  while (G < 32 && (uint64_t)p.nb * G * 2 <= slots && (uint64_t)G * 2 <= (avg + 1) / 2) G *= 2;
  if (tune.acc_G >= 1 && tune.acc_G <= 32 && (tune.acc_G & (tune.acc_G - 1)) == 0) G = (uint32_t)tune.acc_G;
  while (G > 1 && (uint64_t)p.nb * G > p.acc_threads) G /= 2;   // lane sums may pass through the partial-sum buffer
  if (avg / G > 256) return;   // huge buckets without pre-reduction (narrow forced windows): chunks balance better
  p.acc_G = G;
  const uint64_t cap = 4 * avg + 32 * G;
  p.acc_cap = cap > 0x7fffffffu ? 0x7fffffffu : (uint32_t)cap;
}

// slots needed for the partial sums of all fix-up levels
inline size_t msm_partial_slots(const MsmPlan& p) {
  size_t total = 0;
  uint32_t count = p.acc_threads, level = 0;
  for (;;) { total += count; if (count <= 1) break; uint32_t fan = fix_fan(level++); count = (count + fan - 1) / fan; }
  return total + 1;
}

// threads the batched-affine kernel keeps resident: 3 blocks of 128 per SM (ZK_MIN_BLOCKS in tu_g1_batch.cu), i.e.
// 1.5 x acc_slots; tune.batch_blocks overrides the 3 for A/B builds of that translation unit
inline uint64_t msm_batch_slots(const MsmPlan& p, const MsmTuning& tune) {
  const uint64_t blocks = tune.batch_blocks >= 1 && tune.batch_blocks <= 8 ? (uint64_t)tune.batch_blocks : 3;
  return (uint64_t)p.acc_slots * blocks / 2;
}

// Rounds of batched-affine pre-reduction for a set with precomputed slabs on a real device (measured on B200,
// profiles/): an affine addition sharing its inversion costs ~0.29 ns against 0.35 ns for the XYZZ mixed addition
// while a round keeps every resident thread busy with >= 8 additions per inversion; a smaller round is bound by
// the latency of its one inversion per thread (~40 us) and AccumulateBuckets is the better tool.  Rounds stop at
// ~6 items per bucket (measured: 2^22 20.6 -> 20.2 ms, 2^24 74.3 -> 72.8 ms against stopping at 12; 2^20 unchanged).
// The scratch (~250 bytes per pair) is capped.
inline uint32_t msm_default_batch_rounds(const MsmPlan& p, const MsmTuning& tune = MsmTuning()) {
  if (!p.precomp || !p.acc_slots) return 0;
  if ((uint64_t)p.max_entries / 2 * 250 > (40ull << 30)) return 0;
  const uint64_t expected = msm_expected_entries(p), per_bucket = expected / p.nb;
  const uint64_t min_adds = 8 * msm_batch_slots(p, tune);
  uint32_t r = 0;
  const uint64_t min_left = tune.min_left > 0 ? (uint64_t)tune.min_left : 6;
  while (r < 8 && (expected >> (r + 1)) >= min_adds && (per_bucket >> (r + 1)) >= min_left) r++;
  return r < 2 ? 0 : r;
}

// items a batched-affine round can leave (every bucket rounds up), and the prefix-product scratch: a block of
// 128 threads owns 128 * T consecutive slots, so the last block may reach past the item count
inline size_t msm_pre_slots(const MsmPlan& p) { return (size_t)(p.max_entries + 1) / 2 + p.nb + 1; }
inline size_t msm_prefix_slots(const MsmPlan& p) { return msm_pre_slots(p) + (size_t)129 * p.batch_T; }
// additions per thread in a round with about `items` outputs: whole waves of the resident threads, at most
// p.batch_T; below one wave the round's time is one thread's chain (T additions + one inversion), so T shrinks with
// the round (down to 8: the inversion is worth ~6 additions)
inline uint32_t msm_batch_T(const MsmPlan& p, uint64_t items, const MsmTuning& tune = MsmTuning()) {
  if (!p.acc_slots) return p.batch_T;
  const uint64_t slots = msm_batch_slots(p, tune);
  const uint32_t lo = p.batch_T < 8 ? p.batch_T : 8;
  for (uint64_t waves = 1;; waves++) {
    uint64_t t = (items + waves * slots - 1) / (waves * slots);
    if (t <= p.batch_T) return t < lo ? lo : (uint32_t)t;
  }
}

// chunk sums of the bucket reduction (nwin rows of B / K) + scratch rows of the window tree
inline size_t msm_reduced_slots(const MsmPlan& p) {
  size_t chunks = p.B / p.K;
  return (size_t)p.nwin * chunks + (size_t)p.nwin * (chunks / 32 + 1);
}

// device buffers one MSM needs (sizes in elements), all owned by the caller
template <class C> struct MsmBuffers {
  typedef typename C::F F;
  uint32_t* hist_cursor;   // nb
  uint32_t* offsets;       // nb + 1
  uint32_t* segsum;        // ceil(nb / SCAN_SEG)
  Entry* entries;          // max_entries
  XYZZ<F>* bucket_sums;    // nb
  XYZZ<F>* partials;       // msm_partial_slots(): level 0 (one per accumulate thread), then the fix-up levels
  uint32_t* partial_keys;  // same count
  // batched-affine pre-reduction (p.batch_rounds > 0): ping-pong point / offset buffers, prefix scratch
  Affine<F>* pre_pts[2];   // ceil(max_entries / 2) + nb each
  F* pre_prefix;           // 2 * msm_prefix_slots(): prefix products, then the denominators
  uint32_t* pre_off[2];    // nb + 1 each
  uint32_t* pre_cnt;       // nb
  Entry* pre_entries;      // ceil(max_entries / 2) + nb
  XYZZ<F>* reduced;        // msm_reduced_slots(): nwin rows of B / K, then the window tree's scratch
  uint32_t* err;           // 1
  uint32_t* big;           // 1: raised by AccumulateBuckets when a bucket is over the cap (gates the chunked fallback)
};

// The launch sequence.  Exec::launch<Body>(nthreads, args...) runs Body::run(tid, args...) for
// tid in [0, nthreads); Exec::zero(ptr, bytes) clears device memory (both stream-ordered).
template <class C, class Exec>
void msm_launch(Exec& ex, const MsmPlan& p, const MsmTuning& tune, const MsmBuffers<C>& b, const Affine<typename C::F>* points,
                const uint32_t* d_scalars, XYZZ<typename C::F>* out_xyzz, uint32_t* out_affine, uint32_t* out_inf) {
  typedef typename C::F F;
  ex.zero(b.hist_cursor, sizeof(uint32_t) * p.nb);
  ex.zero(b.err, sizeof(uint32_t));
  ex.zero(b.big, sizeof(uint32_t));
  ex.template launch<RecodeCount>(p.n, p, d_scalars, b.hist_cursor, b.err);
  // offsets[0..nb] = exclusive scan of the histogram; the histogram array becomes the scatter cursors
  ex.exclusive_scan(p.nb, b.hist_cursor, b.offsets, b.segsum);
  ex.template launch<Scatter>(p.n, p, d_scalars, b.hist_cursor, b.entries);
  const uint32_t* acc_offsets = b.offsets;
  const Entry* acc_entries = b.entries;
  const Affine<F>* acc_points = points;
  uint64_t bound = p.max_entries;               // worst-case items (sizes the launches)
  uint64_t expected = msm_expected_entries(p);  // items for uniformly random scalars (tunes them)
  if (p.batch_rounds > 0) {
    const uint32_t* off_in = b.offsets;
    const Affine<F>* src = points;
    for (uint32_t r = 0; r < p.batch_rounds; r++) {
      uint32_t* off_out = b.pre_off[r & 1];
      Affine<F>* dst = b.pre_pts[r & 1];
      bool last = r + 1 == p.batch_rounds;
      bound = (bound + 1) / 2 + p.nb;                 // upper bound of this round's outputs
      expected = (expected + 1) / 2 + p.nb / 2;
      MsmPlan pr = p;
      pr.batch_T = msm_batch_T(p, expected, tune);
      const uint32_t threads = (uint32_t)((bound + pr.batch_T - 1) / pr.batch_T);
      ex.scan_pair_counts(p.nb, off_in, b.pre_cnt, off_out, b.segsum);   // off_out = scan of ceil(size_b / 2)
      if (r == 0)
        ex.template launch<BatchedAddRound<C, true>>(threads, pr, off_in, (const uint32_t*)off_out, (const Entry*)b.entries, src, dst,
                                                     b.pre_prefix, msm_prefix_slots(p), last ? b.pre_entries : (Entry*)nullptr);
      else
        ex.template launch<BatchedAddRound<C, false>>(threads, pr, off_in, (const uint32_t*)off_out, (const Entry*)nullptr, src, dst,
                                                      b.pre_prefix, msm_prefix_slots(p), last ? b.pre_entries : (Entry*)nullptr);
      off_in = off_out;
      src = dst;
    }
    acc_offsets = off_in;
    acc_entries = b.pre_entries;
    acc_points = src;
  }
  MsmPlan pa = p;
  if (p.batch_rounds > 0) {
    pa.precomp = 0;   // the pre-reduced items are addressed directly
    // far fewer items are left: shorter chunks keep about two waves of threads busy (never more threads than
    // the partial-sum buffers were sized for)
    uint64_t l = p.acc_slots ? bound / (2ull * p.acc_slots) : p.L;
    uint64_t lmin = (bound * p.L + p.max_entries - 1) / p.max_entries;
    if (l < 4) l = 4;
    if (l < lmin) l = lmin;
    if (l > p.L) l = p.L;
    pa.L = (uint32_t)l;
    pa.acc_threads = (uint32_t)((bound + pa.L - 1) / pa.L);
    if (pa.acc_threads > p.acc_threads) { pa.L = p.L; pa.acc_threads = p.acc_threads; }
  }
  // bucket sums: one launch with acc_G lanes per bucket; the chunked accumulation and its fix-up tree stay as the
  // distribution-independent fallback, gated on the flag a bucket over the cap raises (launches that return at once
  // otherwise), or as the only path when no lane count fits
  msm_pick_bucket_acc(pa, expected, tune);
  const uint32_t* gate = nullptr;   // gates the chunked accumulation (nullptr: always runs)
  if (pa.acc_G) {
    ex.template accumulate_buckets<C>(pa, acc_offsets, acc_entries, acc_points, p.batch_rounds > 0 ? 1u : 0u, b.bucket_sums,
                                      b.partials, b.big);   // (the partial-sum buffer doubles as the lane-sum scratch)
    gate = b.big;
  } else {
    ex.template launch<BucketSizeCheck>(p.nb, p.nb, acc_offsets, 64u * pa.L, b.big);
  }
  // (gated launches go out on a capped grid: those that find the gate closed cost a few hundred blocks each)
  const uint32_t cap_blocks = p.acc_slots ? p.acc_slots / 128 : 296;
  ex.template launch_capped<Accumulate<C>>(gate ? cap_blocks : 0xffffffffu, pa.acc_threads, pa, acc_offsets, acc_entries, acc_points,
                                           b.bucket_sums, b.partials, b.partial_keys, gate);
  if (!pa.acc_G) {   // chunked path as the main path: one-launch fix-up unless a bucket is over the cap
    ex.template launch<FixupDirect<C>>(p.nb, p.nb, pa.L, acc_offsets, (const XYZZ<F>*)b.partials, b.bucket_sums, (const uint32_t*)b.big);
    gate = b.big;    // the tree below runs only when FixupDirect stood back
  }
  {  // fix-up tree over the per-chunk partial sums: level l reads region l, writes region l+1
    uint32_t count = pa.acc_threads, level = 0;
    XYZZ<F>* pin = b.partials;
    uint32_t* kin = b.partial_keys;
    while (count > 1) {
      uint32_t fan = fix_fan_fallback(level++), next = (count + fan - 1) / fan;   // (always gated since FixupDirect exists)
      ex.template launch_capped<FixupLevel<C>>(cap_blocks, next, count, fan, (const uint32_t*)kin, (const XYZZ<F>*)pin, kin + count,
                                               pin + count, b.bucket_sums, gate);
      pin += count;
      kin += count;
      count = next;
    }
  }
  uint32_t chunks = p.B / p.K;
  // stages 6 and 7 go through the Exec policy: the CUDA build has block-cooperative versions (coop.cuh), the CPU
  // emulation runs the per-thread bodies BucketReduce / PairSum; both give the same points
  uint32_t m = ex.template bucket_reduce<C>(p, acc_offsets, (const XYZZ<F>*)b.bucket_sums, b.reduced);
  auto rows = ex.template window_tree<C>(p.nwin, chunks, m, b.reduced, b.reduced + (size_t)p.nwin * chunks);
  ex.template finish<C>(p.nwin, rows.pitch, p.c, rows.arr, out_xyzz, out_affine, out_inf);
}

}  // namespace zk
