// TEST INFRASTRUCTURE ONLY: runs the identical MSM pipeline of csrc/msm.cuh (same bodies, same
// launch sequence) on the CPU, one logical thread after another, with ptx.cuh's instruction
// emulation.  Lets `pytest -m "not gpu"` check the pipeline logic (recoding, counting sort,
// chunked accumulation, fix-up, reduction, Horner, precomputed slabs) against the oracle on a
// machine with no GPU.  Never linked into, nor reachable from, the product library.
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../../zk-toolkit_b200/csrc/msm.cuh"
#include "../../zk-toolkit_b200/csrc/fr_ops.cuh"
#include "../../zk-toolkit_b200/csrc/fr_ntt.cuh"

using namespace zk;

struct HostExec {
  int launches = 0;
  template <class Body, class... Args>
  void launch(uint32_t nthreads, Args... args) {
    for (uint32_t t = 0; t < nthreads; t++) Body::run(t, args...);
    launches++;
  }
  template <class Body, class... Args> void launch_capped(uint32_t, uint32_t nthreads, Args... args) { launch<Body>(nthreads, args...); }
  void zero(void* p, size_t bytes) { memset(p, 0, bytes); }
  template <class C> void accumulate_buckets(const MsmPlan& p, const uint32_t* offsets, const Entry* entries,
                                             const Affine<typename C::F>* points, uint32_t direct,
                                             XYZZ<typename C::F>* bucket_sums, XYZZ<typename C::F>*, uint32_t* big) {
    launch<AccumulateBucketsRef<C>>(p.nb, p, offsets, entries, points, direct, bucket_sums, big);
  }
  template <class C> uint32_t bucket_reduce(const MsmPlan& p, const uint32_t* offsets, const XYZZ<typename C::F>* buckets,
                                            XYZZ<typename C::F>* out) {
    launch<BucketReduce<C>>(p.nwin * (p.B / p.K), p, offsets, buckets, out);
    return p.B / p.K;
  }
  template <class Pt> struct Rows { const Pt* arr; uint32_t pitch; };
  template <class C> Rows<XYZZ<typename C::F>> window_tree(uint32_t nwin, uint32_t pitch, uint32_t m, XYZZ<typename C::F>* arr,
                                                          XYZZ<typename C::F>*) {
    while (m > 1) {
      uint32_t half = (m + 1) / 2;
      launch<PairSum<C>>(nwin * half, nwin, pitch, m, half, arr);
      m = half;
    }
    return {arr, pitch};
  }
  template <class C> void finish(uint32_t nwin, uint32_t pitch, uint32_t c, const XYZZ<typename C::F>* arr,
                                 XYZZ<typename C::F>* out_xyzz, uint32_t* out_affine, uint32_t* out_inf) {
    launch<Finish<C>>(1u, nwin, pitch, c, arr, out_xyzz, out_affine, out_inf);
  }
  bool ntt_fused(bool, Fr*, uint32_t, uint32_t, uint32_t, const Fr*, uint32_t) { return false; }   // per-stage bodies
  void scan_pair_counts(uint32_t nb, const uint32_t* off_in, uint32_t* cnt, uint32_t* off_out, uint32_t* segsum) {
    launch<PairCount>(nb, nb, off_in, cnt);      // the unfused formulation; the CUDA build forms the counts inside its scan
    exclusive_scan(nb, cnt, off_out, segsum);
  }
  void exclusive_scan(uint32_t nb, uint32_t* hist_cursor, uint32_t* offsets, uint32_t* segsum) {
    uint32_t nseg = (nb + SCAN_SEG - 1) / SCAN_SEG;
    launch<ScanLocal>(nseg, nb, (const uint32_t*)hist_cursor, segsum);
    launch<ScanTop>(1u, nseg, nb, segsum, offsets);
    launch<ScanApply>(nseg, nb, hist_cursor, (const uint32_t*)segsum, offsets);
  }
};

static uint32_t g_last_fallback = 0;   // 1 when the last emu_msm took the chunked fallback (a bucket over the cap)

template <class C>
static int emu_msm(const uint32_t* xy, const uint8_t* inf, const uint32_t* scalars, uint32_t n, uint32_t npts,
                   uint32_t c, int precomp, uint32_t L, uint32_t K, uint32_t* out_xy, uint32_t* out_inf,
                   uint32_t* out_xyzz, uint32_t rank = 0, uint32_t world = 1) {
  typedef typename C::F F;
  HostExec ex;
  bool half = (precomp & 2) != 0;
  precomp &= 1;
  if (c == 0) c = msm_pick_c(precomp ? npts : n, precomp != 0, half);
  uint32_t W = msm_windows(c, half);
  std::vector<Affine<F>> pts((size_t)npts * (precomp ? W : 1) + 1);
  ex.launch<LoadPoints<C>>(npts, npts, xy, inf, pts.data());
  if (precomp) ex.launch<PrecomputeSlabs<C>>(npts, npts, npts, c, W, msm_wide_windows(c, W, half, true), pts.data());
  if (n == 0) { *out_inf = 1; memset(out_xy, 0, 4 * C::AFF_LIMBS); return 0; }
  const MsmTuning tune = MsmTuning::from_env();   // test infrastructure: the switches are read per call here
  MsmPlan p = msm_plan(n, c, precomp != 0, npts, half, false, 0, tune, rank, world);
  if (L) { p.L = L; p.acc_threads = (p.max_entries + L - 1) / L; }
  if (K) p.K = K;
  std::vector<uint32_t> hist(p.nb), offsets(p.nb + 1), segsum((p.nb + SCAN_SEG - 1) / SCAN_SEG + 1), err(1), big(1);
  std::vector<Entry> entries(p.max_entries + 1);
  std::vector<XYZZ<F>> buckets(p.nb), partials(msm_partial_slots(p)), reduced(msm_reduced_slots(p));
  std::vector<uint32_t> pkeys(msm_partial_slots(p), 0x12345678u);
  // poison what the pipeline must overwrite before reading
  memset(buckets.data(), 0xAB, sizeof(XYZZ<F>) * buckets.size());
  memset(partials.data(), 0xCD, sizeof(XYZZ<F>) * partials.size());
  const size_t pre_n = msm_pre_slots(p);
  std::vector<Affine<F>> pre_a(p.batch_rounds ? pre_n : 1), pre_b(p.batch_rounds ? pre_n : 1);
  std::vector<F> pre_prefix(p.batch_rounds ? 2 * msm_prefix_slots(p) : 1);
  std::vector<uint32_t> pre_off_a(p.nb + 1), pre_off_b(p.nb + 1), pre_cnt(p.nb + 1);
  std::vector<Entry> pre_entries(p.batch_rounds ? pre_n : 1);
  MsmBuffers<C> b;
  b.pre_pts[0] = pre_a.data(); b.pre_pts[1] = pre_b.data(); b.pre_prefix = pre_prefix.data();
  b.pre_off[0] = pre_off_a.data(); b.pre_off[1] = pre_off_b.data(); b.pre_cnt = pre_cnt.data(); b.pre_entries = pre_entries.data();
  b.hist_cursor = hist.data(); b.offsets = offsets.data(); b.segsum = segsum.data(); b.entries = entries.data();
  b.bucket_sums = buckets.data(); b.partials = partials.data(); b.reduced = reduced.data(); b.err = err.data(); b.big = big.data(); b.partial_keys = pkeys.data();
  XYZZ<F> xyzz;
  msm_launch<C>(ex, p, tune, b, (const Affine<F>*)pts.data(), scalars, out_xyzz ? &xyzz : (XYZZ<F>*)nullptr,
                out_xyzz ? (uint32_t*)nullptr : out_xy, out_xyzz ? (uint32_t*)nullptr : out_inf);
  if (out_xyzz) {
    if (err[0]) ex.launch<PoisonPartial<C>>(1u, (const uint32_t*)err.data(), &xyzz);
    memcpy(out_xyzz, &xyzz, sizeof(xyzz));
  }
  g_last_fallback = big[0];
  return err[0] ? -3 : 0;
}

extern "C" {
int emu_g1_msm(const uint32_t* xy, const uint8_t* inf, const uint32_t* scalars, uint32_t n, uint32_t npts, uint32_t c,
               int precomp, uint32_t L, uint32_t K, uint32_t* out_xy, uint32_t* out_inf) {
  return emu_msm<G1>(xy, inf, scalars, n, npts, c, precomp, L, K, out_xy, out_inf, nullptr);
}
int emu_g2_msm(const uint32_t* xy, const uint8_t* inf, const uint32_t* scalars, uint32_t n, uint32_t npts, uint32_t c,
               int precomp, uint32_t L, uint32_t K, uint32_t* out_xy, uint32_t* out_inf) {
  return emu_msm<G2>(xy, inf, scalars, n, npts, c, precomp, L, K, out_xy, out_inf, nullptr);
}
// k shards -> k partials -> CombinePartials (the multi-GPU path, one "rank" after another)
int emu_g1_msm_sharded(const uint32_t* xy, const uint32_t* scalars, uint32_t n, uint32_t k, uint32_t* out_xy,
                       uint32_t* out_inf) {
  std::vector<XYZZ<Fp>> parts(k);
  for (uint32_t r = 0; r < k; r++) {
    uint32_t lo = (uint64_t)n * r / k, hi = (uint64_t)n * (r + 1) / k;
    if (hi == lo) { memset(&parts[r], 0, sizeof(XYZZ<Fp>)); continue; }
    uint32_t dummy[24], dinf;
    int rc = emu_msm<G1>(xy + 24 * (size_t)lo, nullptr, scalars + 8 * (size_t)lo, hi - lo, hi - lo, 0, 0, 0, 0, dummy, &dinf,
                         (uint32_t*)&parts[r]);
    if (rc) return rc;
  }
  HostExec ex;
  uint32_t err = 0;
  ex.launch<CombinePartials<G1>>(1u, k, (const XYZZ<Fp>*)parts.data(), (uint32_t)(sizeof(XYZZ<Fp>) / 4), out_xy, out_inf, &err);
  return err ? -3 : 0;
}
// bucket-range split: `world` ranks each see all n scalars and the whole precomputed set, own 1/world of the buckets
int emu_g1_msm_range(const uint32_t* xy, const uint32_t* scalars, uint32_t n, uint32_t world, uint32_t c, int half, uint32_t* out_xy,
                     uint32_t* out_inf) {
  std::vector<XYZZ<Fp>> parts(world);
  for (uint32_t r = 0; r < world; r++) {
    uint32_t dummy[24], dinf;
    int rc = emu_msm<G1>(xy, nullptr, scalars, n, n, c, 1 | (half ? 2 : 0), 0, 0, dummy, &dinf, (uint32_t*)&parts[r], r, world);
    if (rc) return rc;
  }
  HostExec ex;
  uint32_t err = 0;
  ex.launch<CombinePartials<G1>>(1u, world, (const XYZZ<Fp>*)parts.data(), (uint32_t)(sizeof(XYZZ<Fp>) / 4), out_xy, out_inf, &err);
  return err ? -3 : 0;
}
uint32_t emu_last_fallback() { return g_last_fallback; }
// one rank's share under the bucket-range split, as the 48-word blob (zkmsm_g1_msm_partial_range)
int emu_g1_msm_partial_range(const uint32_t* xy, const uint32_t* scalars, uint32_t n, uint32_t c, int half, uint32_t rank,
                             uint32_t world, uint32_t* out_partial) {
  uint32_t dummy[24], dinf;
  return emu_msm<G1>(xy, nullptr, scalars, n, n, c, 1 | (half ? 2 : 0), 0, 0, dummy, &dinf, out_partial, rank, world);
}
// one rank's share: partial point as the opaque 48-word blob of the C ABI, and the combine step
int emu_g1_msm_partial(const uint32_t* xy, const uint32_t* scalars, uint32_t n, uint32_t* out_partial) {
  if (n == 0) { memset(out_partial, 0, sizeof(XYZZ<Fp>)); return 0; }
  uint32_t dummy[24], dinf;
  return emu_msm<G1>(xy, nullptr, scalars, n, n, 0, 0, 0, 0, dummy, &dinf, out_partial);
}
int emu_g1_combine(const uint32_t* partials, uint32_t k, uint32_t* out_xy, uint32_t* out_inf) {
  HostExec ex;
  uint32_t err = 0;
  ex.launch<CombinePartials<G1>>(1u, k, (const XYZZ<Fp>*)partials, (uint32_t)(sizeof(XYZZ<Fp>) / 4), out_xy, out_inf, &err);
  return err ? -3 : 0;
}
// the same over partials that sit stride_words apart (the per-rank blobs of a distributed proof, zkmsm_groth16_combine)
int emu_g1_combine_strided(const uint32_t* partials, uint32_t k, uint32_t stride_words, uint32_t* out_xy, uint32_t* out_inf) {
  HostExec ex;
  uint32_t err = 0;
  ex.launch<CombinePartials<G1>>(1u, k, (const XYZZ<Fp>*)partials, stride_words, out_xy, out_inf, &err);
  return err ? -3 : 0;
}
// `base * k_i` for a vector of raw 256-bit scalars (FixedBaseMul path)
int emu_g1_mul_base(const uint32_t* base_xy, const uint32_t* scalars, uint32_t n, uint32_t* out_xy, uint8_t* out_inf) {
  HostExec ex;
  std::vector<XYZZ<Fp>> chain(256);
  std::vector<Affine<Fp>> table(256), out(n + 1);
  ex.launch<BaseTableChain<G1>>(1u, base_xy, chain.data());
  ex.launch<BaseTableAffine<G1>>(256u, (const XYZZ<Fp>*)chain.data(), table.data());
  ex.launch<FixedBaseMul<G1>>(n, n, scalars, (const Affine<Fp>*)table.data(), out.data());
  ex.launch<StorePoints<G1>>(n, n, (const Affine<Fp>*)out.data(), out_xy, out_inf);
  return 0;
}
int emu_fr_aggregate(const uint32_t* polys, uint32_t n_wires, uint32_t n, const uint32_t* wires, uint32_t* out) {
  HostExec ex;
  std::vector<Fr> wm(n_wires + 1);
  ex.launch<FrToMont>(n_wires, n_wires, wires, wm.data());
  ex.launch<FrAggregate>(n, n_wires, n, (const Fr*)wm.data(), polys, out);
  return 0;
}
// quotient polynomial: same launch sequence as zkmsm_fr_quotient
int emu_fr_quotient(const uint32_t* u, const uint32_t* v, const uint32_t* w, uint32_t n, uint32_t* h_out, uint32_t* nonzero_rem) {
  HostExec ex;
  std::vector<Fr> du(n), dv(n), dw(n), dp(2 * n), t0(n + 2), t1(n + 2), dh(n);
  ex.launch<FrVecToMont>(n, n, n, u, du.data());
  ex.launch<FrVecToMont>(n, n, n, v, dv.data());
  ex.launch<FrVecToMont>(n, n, n, w, dw.data());
  ex.launch<FrPolyMulSub>(2 * n - 1, n, (const Fr*)du.data(), (const Fr*)dv.data(), (const Fr*)dw.data(), dp.data());
  for (auto& x : t0) fset_zero(x);
  for (auto& x : t1) fset_zero(x);
  fset_one(t0[0]);
  Fr* t_old = t0.data();
  Fr* t_new = t1.data();
  for (uint32_t k = 1; k <= n; k++) {
    uint32_t kc[8] = {k, 0, 0, 0, 0, 0, 0, 0};
    Fr km;
    fto_mont(km, kc);
    ex.launch<FrTStep>(k + 1, k, km, (const Fr*)t_old, t_new);
    Fr* tmp = t_old; t_old = t_new; t_new = tmp;
  }
  for (int d = (int)n - 2; d >= 0; d--) ex.launch<FrDivStep>(n, n, (uint32_t)d, (const Fr*)t_old, dp.data(), dh.data());
  *nonzero_rem = 0;
  ex.launch<FrQuotientOut>(n, n, (const Fr*)dh.data(), (const Fr*)dp.data(), h_out, nonzero_rem);
  return 0;
}
// the transform-based quotient (fr_ntt.cuh): same pipeline as zkmsm_fr_quotient's n >= 32 path
int emu_fr_quotient_ntt(const uint32_t* u, const uint32_t* v, const uint32_t* w, uint32_t n, uint32_t* h_out, uint32_t* nonzero_rem,
                        uint32_t* t_out) {
  HostExec ex;
  const FrNttPlan p = FrNttPlan::make(n);
  std::vector<Fr> tables(fr_ntt_table_elems(p)), scratch(5 * (size_t)p.m);
  FrNttTables tb;
  fr_ntt_tables_at(tb, p, tables.data());
  fr_quotient_setup(ex, p, tb, scratch.data());
  if (t_out)
    for (uint32_t i = 0; i <= n; i++) ffrom_mont(t_out + (size_t)i * 8, tb.t[i]);
  *nonzero_rem = 0;
  fr_quotient_run(ex, p, tb, u, v, w, scratch.data(), h_out, nonzero_rem);
  return ex.launches;
}
// host-side policies of the launch plan on a device with `acc_slots` resident accumulate threads:
// out = {c, L, default batched-affine rounds, additions per thread of the first round, its thread count, K, coop}
void emu_policy(uint32_t n, uint32_t c, int precomp, int half, uint32_t acc_slots, uint32_t* out) {
  if (c == 0) c = msm_pick_c(n, precomp != 0, half != 0);
  const uint32_t world = out[7] ? out[7] : 1;   // in: ranks of a bucket-range split (0 / 1 = none)
  MsmPlan p = msm_plan(n, c, precomp != 0, n, half != 0, true, acc_slots, MsmTuning(), 0, world);
  uint32_t rounds = msm_default_batch_rounds(p);
  uint64_t items = (msm_expected_entries(p) + 1) / 2 + p.nb / 2;
  uint32_t T = msm_batch_T(p, items);
  uint64_t left = msm_expected_entries(p);
  for (uint32_t r = 0; r < rounds; r++) left = (left + 1) / 2 + p.nb / 2;
  msm_pick_bucket_acc(p, left, MsmTuning());
  out[0] = c; out[1] = p.L; out[2] = rounds; out[3] = T; out[4] = (uint32_t)((items + T - 1) / T); out[5] = p.K; out[6] = p.coop;
  out[7] = p.acc_G; out[8] = p.acc_cap; out[9] = p.nb;
}
// ZKMSM_CHECK_SUBGROUP: number of points with r P != AtInfinity
uint32_t emu_g1_subgroup_check(const uint32_t* xy, const uint8_t* inf, uint32_t n) {
  HostExec ex;
  std::vector<Affine<Fp>> pts(n + 1);
  ex.launch<LoadPoints<G1>>(n, n, xy, inf, pts.data());
  uint32_t bad = 0;
  ex.launch<SubgroupCheck<G1>>(n, n, (const Affine<Fp>*)pts.data(), &bad);
  return bad;
}
// n * windows must stay below 2^32 sorted pairs (the kernels' positions are 32-bit): the product rejects the rest
int emu_fits(uint64_t n, uint32_t c, int half) { return msm_fits(n, c, half != 0) ? 1 : 0; }
void emu_plan(uint32_t n, uint32_t c, int precomp, uint32_t* out) {
  if (c == 0) c = msm_pick_c(n, precomp != 0);
  MsmPlan p = msm_plan(n, c, precomp != 0, n);
  memcpy(out, &p, sizeof(p));
}
}
