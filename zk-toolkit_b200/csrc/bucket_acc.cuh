// Bucket sums in one launch: acc_G lanes per bucket (BucketAccLane, msm.cuh), the lane sums meeting in a register
// shuffle tree.  Replaces the chunked accumulation + fix-up tree whenever no bucket is over the plan's cap; see
// AccumulateBucketsRef in msm.cuh for the reference formulation the CPU emulation runs.
#pragma once
#include <cuda_runtime.h>
#include "msm.cuh"

namespace zk {

template <class F> __device__ __forceinline__ void shfl_down_xyzz(XYZZ<F>& dst, const XYZZ<F>& src, uint32_t delta, uint32_t width) {
  const uint32_t* s = reinterpret_cast<const uint32_t*>(&src);
  uint32_t* d = reinterpret_cast<uint32_t*>(&dst);
#pragma unroll
  for (int i = 0; i < (int)(sizeof(XYZZ<F>) / 4); i++) d[i] = __shfl_down_sync(0xffffffffu, s[i], delta, width);
}

// SPLIT (G2): two 96-word points do not fit the register file next to the general addition's temporaries, and
// compiling the lane tree into this kernel makes the accumulation loop spill (measured 8.4 ms against 4.6 ms for
// the same mixed additions at 2^18).  The lanes then leave their sums in `lane_sums` (one slot per thread) and
// BucketLaneSum (msm.cuh), a second small launch, adds the G sums of every bucket.  Otherwise the lane sums meet
// in a register-shuffle tree.
template <class C, bool SPLIT>
__global__ void __launch_bounds__(128, 2) bucket_acc_kernel(MsmPlan p, const uint32_t* offsets, const Entry* entries,
                                                            const Affine<typename C::F>* points, uint32_t direct,
                                                            XYZZ<typename C::F>* bucket_sums, XYZZ<typename C::F>* lane_sums,
                                                            uint32_t* big) {
  typedef typename C::F F;
  const uint32_t G = p.acc_G, tid = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t b = tid / G, g = tid % G;
  const bool valid = b < p.nb;
  const uint32_t s = valid ? offsets[b] : 0, e = valid ? offsets[b + 1] : 0;
  XYZZ<F> acc;
  const bool ok = BucketAccLane<C>::run(acc, s, e, g, G, p.acc_cap, entries, points, direct);
  if constexpr (SPLIT) {
    if (!valid || s == e) return;
    if (!ok) { if (g == 0) atomicOr(big, 1u); return; }
    if (G > 1) lane_sums[tid] = acc; else bucket_sums[b] = acc;
  } else {
    for (uint32_t d = G >> 1; d >= 1; d >>= 1) {   // all 32 lanes take part in every shuffle (G divides 32)
      XYZZ<F> other;
      shfl_down_xyzz(other, acc, d, G);
      if (g < d) xyzz_add(acc, other);
    }
    if (g != 0 || s == e) return;
    if (!ok) { atomicOr(big, 1u); return; }
    bucket_sums[b] = acc;
  }
}

template <class C>
cudaError_t bucket_acc_launch(cudaStream_t st, const MsmPlan& p, const uint32_t* offsets, const Entry* entries,
                              const Affine<typename C::F>* points, uint32_t direct, XYZZ<typename C::F>* bucket_sums,
                              XYZZ<typename C::F>* lane_sums, uint32_t* big) {
  const uint64_t threads = (uint64_t)p.nb * p.acc_G;
  constexpr bool kSplit = sizeof(typename C::F) > 48;
  bucket_acc_kernel<C, kSplit><<<(unsigned)((threads + 127) / 128), 128, 0, st>>>(p, offsets, entries, points, direct, bucket_sums,
                                                                                lane_sums, big);
  return cudaGetLastError();
}

}  // namespace zk
