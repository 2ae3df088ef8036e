// Kernel launch plumbing.  A pipeline stage is a "body" struct with
//   static void run(uint32_t tid, Args...);
// launched over nthreads logical threads by Launch<Body>::go.  The argument types are taken
// from Body::run's own signature, so the only key of a launcher is the body type: translation
// units that own a kernel define ZK_DEFINE_LAUNCH and explicitly instantiate
//   template struct zk::LaunchBase<Body, decltype(Body::run)>;
// every other translation unit only sees the declaration and links against it.  That keeps the
// heavy field-arithmetic kernels in their own files, compiled in parallel.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <type_traits>

namespace zk {

static constexpr uint32_t kBlock = 128;

#ifndef ZK_MIN_BLOCKS
#define ZK_MIN_BLOCKS 1   // a translation unit may ask ptxas to fit this many 128-thread blocks per SM
#endif

template <class Body, class... A>
__global__ void __launch_bounds__(kBlock, ZK_MIN_BLOCKS) body_kernel(uint32_t nthreads, A... a) {
  uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (tid < nthreads) Body::run(tid, a...);
}

// same, on a grid of at most max_blocks blocks that strides over the logical threads: what the gated fallback
// launches use (a launch that returns at its gate then costs ~300 blocks instead of thousands)
template <class Body, class... A>
__global__ void __launch_bounds__(kBlock, ZK_MIN_BLOCKS) body_kernel_strided(uint32_t nthreads, A... a) {
  for (uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x; tid < nthreads; tid += gridDim.x * blockDim.x) Body::run(tid, a...);
}

template <class Body, class Sig> struct LaunchBase;
template <class Body, class... A> struct LaunchBase<Body, void(uint32_t, A...)> {
  static cudaError_t go(cudaStream_t st, uint32_t nthreads, A... a)
#ifdef ZK_DEFINE_LAUNCH
  {
    body_kernel<Body, A...><<<(nthreads + kBlock - 1) / kBlock, kBlock, 0, st>>>(nthreads, a...);
    return cudaGetLastError();
  }
#else
      ;
#endif
};
// the strided form is its own launcher so that only the bodies that use it (ZK_INSTANTIATE_KERNEL_STRIDED) compile it
template <class Body, class Sig> struct LaunchStridedBase;
template <class Body, class... A> struct LaunchStridedBase<Body, void(uint32_t, A...)> {
  static cudaError_t go(cudaStream_t st, uint32_t max_blocks, uint32_t nthreads, A... a)
#ifdef ZK_DEFINE_LAUNCH
  {
    uint32_t blocks = (nthreads + kBlock - 1) / kBlock;
    body_kernel_strided<Body, A...><<<blocks < max_blocks ? blocks : max_blocks, kBlock, 0, st>>>(nthreads, a...);
    return cudaGetLastError();
  }
#else
      ;
#endif
};
template <class Body> struct LaunchStrided : LaunchStridedBase<Body, decltype(Body::run)> {};
template <class Body> struct Launch : LaunchBase<Body, decltype(Body::run)> {};

#define ZK_INSTANTIATE_KERNEL(...) template struct zk::LaunchBase<__VA_ARGS__, decltype(__VA_ARGS__::run)>
#define ZK_INSTANTIATE_KERNEL_STRIDED(...) template struct zk::LaunchStridedBase<__VA_ARGS__, decltype(__VA_ARGS__::run)>

// Exec policy for msm_launch (msm.cuh): stream-ordered CUDA launches
// block-cooperative exclusive scan (tu_sort.cu): offsets[0..n] = scan(hist), hist <- offsets (cursors)
cudaError_t zk_exclusive_scan(cudaStream_t st, uint32_t n, uint32_t* hist_cursor, uint32_t* offsets, uint32_t* blocksums);
// off_out[0..n] = scan of the per-bucket output counts ceil(size / 2) of a batched-affine round (sizes from off_in)
cudaError_t zk_exclusive_scan_pairs(cudaStream_t st, uint32_t n, const uint32_t* off_in, uint32_t* off_out, uint32_t* blocksums);
int zk_scan_launches(uint32_t n);   // kernels one scan of n elements launches (1 up to 8192 elements, else 3)

// block-cooperative G1 tail kernels (tu_g1_coop.cu)
struct MsmPlan;
struct FqCfg;
template <class Cfg> struct Mont;
template <class F> struct XYZZ;
cudaError_t zk_coop_bucket_reduce_g1(cudaStream_t st, const MsmPlan& p, const uint32_t* offsets,
                                     const XYZZ<Mont<FqCfg>>* buckets, XYZZ<Mont<FqCfg>>* out, int tree);
cudaError_t zk_coop_row_sum_g1(cudaStream_t st, uint32_t nwin, uint32_t pitch_in, uint32_t m, uint32_t per_block,
                               const XYZZ<Mont<FqCfg>>* in, uint32_t pitch_out, XYZZ<Mont<FqCfg>>* out);
cudaError_t zk_coop_finish_g1(cudaStream_t st, uint32_t nwin, uint32_t pitch, uint32_t c, const XYZZ<Mont<FqCfg>>* arr,
                              XYZZ<Mont<FqCfg>>* out_xyzz, uint32_t* out_affine, uint32_t* out_inf);
cudaError_t zk_coop_combine_g1(cudaStream_t st, uint32_t k, const XYZZ<Mont<FqCfg>>* parts, uint32_t stride_words, uint32_t* out_affine,
                               uint32_t* out_inf, uint32_t* err);
template <class C> struct Finish;
struct Fp2;
struct Entry;
template <class F> struct Affine;
cudaError_t zk_bucket_acc_g1(cudaStream_t st, const MsmPlan& p, const uint32_t* offsets, const Entry* entries,
                             const Affine<Mont<FqCfg>>* points, uint32_t direct, XYZZ<Mont<FqCfg>>* bucket_sums,
                             XYZZ<Mont<FqCfg>>* lane_sums, uint32_t* big);
cudaError_t zk_bucket_acc_g2(cudaStream_t st, const MsmPlan& p, const uint32_t* offsets, const Entry* entries,
                             const Affine<Fp2>* points, uint32_t direct, XYZZ<Fp2>* bucket_sums, XYZZ<Fp2>* lane_sums, uint32_t* big);
template <class C> struct BucketLaneSum;
// per-device opt-in to > 48 KB of dynamic shared memory for every kernel that needs it (once per device, before the
// first launch there; zkmsm_create calls it for the context's device)
cudaError_t zk_opt_in_shared_memory_coop_g1();
cudaError_t zk_opt_in_shared_memory_coop_g2();
cudaError_t zk_opt_in_shared_memory_ntt();
cudaError_t zk_coop_bucket_reduce_g2(cudaStream_t st, const MsmPlan& p, const uint32_t* offsets, const XYZZ<Fp2>* buckets,
                                     XYZZ<Fp2>* out, int tree);
cudaError_t zk_coop_combine_g2(cudaStream_t st, uint32_t k, const XYZZ<Fp2>* parts, uint32_t stride_words, uint32_t* out_affine,
                               uint32_t* out_inf, uint32_t* err);
cudaError_t zk_coop_row_sum_g2(cudaStream_t st, uint32_t nwin, uint32_t pitch_in, uint32_t m, uint32_t per_block,
                               const XYZZ<Fp2>* in, uint32_t pitch_out, XYZZ<Fp2>* out);
struct FrCfg;
cudaError_t zk_ntt_fused(cudaStream_t st, bool inverse, Mont<FrCfg>* a, uint32_t total, uint32_t L0, uint32_t k,
                         const Mont<FrCfg>* tw, uint32_t m);
template <class C> struct BucketReduce;
template <class C> struct PairSum;
struct G1;
struct G2;

// optional per-launch timing (zkmsm_profile): CUDA events on the launching stream around every kernel
struct LaunchProfile {
  static constexpr int MAX = 256;
  int n = 0;
  const char* names[MAX];
  uint32_t threads[MAX];
  cudaEvent_t beg[MAX], end[MAX];
  bool have_events = false;
};

template <class Pt> struct RowRef { const Pt* arr; uint32_t pitch; };

struct CudaExec {
  cudaStream_t st;
  int launches;
  cudaError_t err;
  LaunchProfile* prof;
  bool no_coop, ntt_no_fuse;   // cross-check switches (MsmTuning), fixed per context
  explicit CudaExec(cudaStream_t s, LaunchProfile* p = nullptr, bool no_coop_ = false, bool ntt_no_fuse_ = false)
      : st(s), launches(0), err(cudaSuccess), prof(p), no_coop(no_coop_), ntt_no_fuse(ntt_no_fuse_) {}
  template <class Body, class... Args>
  void launch(uint32_t nthreads, Args... args) {
    if (nthreads == 0 || err != cudaSuccess) return;
    int slot = -1;
    if (prof && prof->n < LaunchProfile::MAX) {
      slot = prof->n++;
      prof->names[slot] = Body::name();
      prof->threads[slot] = nthreads;
      cudaEventRecord(prof->beg[slot], st);
    }
    cudaError_t e = Launch<Body>::go(st, nthreads, args...);
    if (slot >= 0) cudaEventRecord(prof->end[slot], st);
    launches++;
    if (e != cudaSuccess) err = e;
  }
  // Body launch on a capped, striding grid (see body_kernel_strided)
  template <class Body, class... Args>
  void launch_capped(uint32_t max_blocks, uint32_t nthreads, Args... args) {
    if (nthreads == 0) return;
    timed(Body::name(), nthreads, 1, [&] { return LaunchStrided<Body>::go(st, max_blocks, nthreads, args...); });
  }
  // profiling bracket around a non-Body launch
  template <class Fn> void timed(const char* name, uint32_t threads, int nlaunch, Fn fn) {
    if (err != cudaSuccess) return;
    int slot = -1;
    if (prof && prof->n < LaunchProfile::MAX) {
      slot = prof->n++;
      prof->names[slot] = name;
      prof->threads[slot] = threads;
      cudaEventRecord(prof->beg[slot], st);
    }
    cudaError_t e = fn();
    if (slot >= 0) cudaEventRecord(prof->end[slot], st);
    launches += nlaunch;
    if (e != cudaSuccess) err = e;
  }
  // bucket sums with acc_G lanes per bucket (tu_g{1,2}_bacc.cu): the lane sums meet in a shuffle tree (G1) or, for G2,
  // are left in lane_sums and added per bucket by a second small launch
  template <class C, class P, class E, class A, class Pt>
  void accumulate_buckets(const P& p, const uint32_t* offsets, const E* entries, const A* points, uint32_t direct, Pt* bucket_sums,
                          Pt* lane_sums, uint32_t* big) {
    const uint32_t threads = p.nb * p.acc_G;
    if (std::is_same<C, G1>::value) {
      timed("accumulate_buckets", threads, 1, [&] { return zk_bucket_acc_g1(st, p, offsets, entries, (const Affine<Mont<FqCfg>>*)points, direct, (XYZZ<Mont<FqCfg>>*)bucket_sums, (XYZZ<Mont<FqCfg>>*)lane_sums, big); });
    } else {
      timed("accumulate_buckets", threads, 1, [&] { return zk_bucket_acc_g2(st, p, offsets, entries, (const Affine<Fp2>*)points, direct, (XYZZ<Fp2>*)bucket_sums, (XYZZ<Fp2>*)lane_sums, big); });
      if (p.acc_G > 1) launch<BucketLaneSum<C>>(p.nb, p, offsets, (const Pt*)lane_sums, bucket_sums, (const uint32_t*)big);
    }
  }
  // stages 6 / 7 of the MSM: block-cooperative kernels (coop.cuh), per-thread bodies as the wide / fallback path.
  // bucket_reduce returns the row length it left per window (row pitch stays B / K): the cooperative kernel also
  // sums the 32 chains of each block when they share a window.
  template <class C, class P, class Pt>
  uint32_t bucket_reduce(const P& p, const uint32_t* offsets, const Pt* buckets, Pt* out) {
    uint32_t chunks = p.B / p.K, chains = p.nwin * chunks;
    if (chains == 0) return chunks;
    if (p.coop) {
      int tree = chunks % 32 == 0 ? 1 : 0;
      if (std::is_same<C, G1>::value)
        timed("bucket_reduce", chains, 1, [&] { return zk_coop_bucket_reduce_g1(st, p, offsets, (const XYZZ<Mont<FqCfg>>*)buckets, (XYZZ<Mont<FqCfg>>*)out, tree); });
      else
        timed("bucket_reduce", chains, 1, [&] { return zk_coop_bucket_reduce_g2(st, p, offsets, (const XYZZ<Fp2>*)buckets, (XYZZ<Fp2>*)out, tree); });
      return tree ? chunks / 32 : chunks;
    }
    launch<BucketReduce<C>>(chains, p, offsets, buckets, out);
    return chunks;
  }
  // Reduces every window's row (m elements, pitch apart) to one point; returns where the sums are (element 0 of
  // each row of the returned buffer).  Wide levels halve in place with one thread per pair; once a level fits
  // the cooperative kernels a block sums up to 128 elements (3 additions per lane, then the lane tree), the
  // levels alternating between arr and scratch (nwin * (pitch / 32 + 1) points).
  template <class C, class Pt>
  RowRef<Pt> window_tree(uint32_t nwin, uint32_t pitch, uint32_t m, Pt* arr, Pt* scratch) {
    const bool coop_ok = !no_coop;
    while (m > 1 && (!coop_ok || (uint64_t)nwin * m > 16384)) {
      uint32_t half = (m + 1) / 2;
      launch<PairSum<C>>(nwin * half, nwin, pitch, m, half, arr);
      m = half;
    }
    Pt* in = arr;
    Pt* out = scratch;
    uint32_t pitch_in = pitch, pitch_out = pitch / 32 + 1;
    while (m > 1) {
      uint32_t per_block = m <= 128 ? m : 128, next = (m + per_block - 1) / per_block;
      if (std::is_same<C, G1>::value)
        timed("row_sum", nwin * next * 32, 1, [&] { return zk_coop_row_sum_g1(st, nwin, pitch_in, m, per_block, (const XYZZ<Mont<FqCfg>>*)in, pitch_out, (XYZZ<Mont<FqCfg>>*)out); });
      else
        timed("row_sum", nwin * next * 32, 1, [&] { return zk_coop_row_sum_g2(st, nwin, pitch_in, m, per_block, (const XYZZ<Fp2>*)in, pitch_out, (XYZZ<Fp2>*)out); });
      Pt* t = in; in = out; out = t;
      uint32_t tp = pitch_in; pitch_in = pitch_out; pitch_out = tp;
      m = next;
    }
    return RowRef<Pt>{in, pitch_in};
  }
  // stage 8: Horner over the windows runs cooperatively for G1 when there is more than one window
  template <class C, class Pt>
  void finish(uint32_t nwin, uint32_t pitch, uint32_t c, const Pt* arr, Pt* out_xyzz, uint32_t* out_affine, uint32_t* out_inf) {
    if (std::is_same<C, G1>::value && nwin > 1 && !no_coop)
      timed("finish", 1, 1, [&] { return zk_coop_finish_g1(st, nwin, pitch, c, (const XYZZ<Mont<FqCfg>>*)arr, (XYZZ<Mont<FqCfg>>*)out_xyzz, out_affine, out_inf); });
    else
      launch<Finish<C>>(1u, nwin, pitch, c, arr, out_xyzz, out_affine, out_inf);
  }
  // several transform stages in one shared-memory pass (tu_fr_ntt.cu); false = shape not supported, launch the stages
  bool ntt_fused(bool inverse, Mont<FrCfg>* a, uint32_t total, uint32_t L0, uint32_t k, const Mont<FrCfg>* tw, uint32_t m) {
    if (err != cudaSuccess) return true;
    if (ntt_no_fuse) return false;
    int slot = -1;
    if (prof && prof->n < LaunchProfile::MAX) {
      slot = prof->n;
      prof->names[slot] = "ntt_fused";
      prof->threads[slot] = total;
      cudaEventRecord(prof->beg[slot], st);
    }
    cudaError_t e = zk_ntt_fused(st, inverse, a, total, L0, k, tw, m);
    if (e == cudaErrorNotSupported) { cudaGetLastError(); return false; }
    if (slot >= 0) { cudaEventRecord(prof->end[slot], st); prof->n++; }
    launches++;
    if (e != cudaSuccess) err = e;
    return true;
  }
  void exclusive_scan(uint32_t n, uint32_t* hist_cursor, uint32_t* offsets, uint32_t* blocksums) {
    timed("exclusive_scan", n, zk_scan_launches(n), [&] { return zk_exclusive_scan(st, n, hist_cursor, offsets, blocksums); });
  }
  // output offsets of a batched-affine round: scan of ceil(size / 2) per bucket, the counts formed inside the scan
  // (cnt_scratch is what the unfused formulation -- PairCount, then a scan -- needs; unused here)
  void scan_pair_counts(uint32_t n, const uint32_t* off_in, uint32_t* /*cnt_scratch*/, uint32_t* off_out, uint32_t* blocksums) {
    timed("scan_pair_counts", n, zk_scan_launches(n), [&] { return zk_exclusive_scan_pairs(st, n, off_in, off_out, blocksums); });
  }
  void zero(void* p, size_t bytes) {
    if (err != cudaSuccess || bytes == 0) return;
    cudaError_t e = cudaMemsetAsync(p, 0, bytes, st);
    if (e != cudaSuccess) err = e;
  }
};

}  // namespace zk
