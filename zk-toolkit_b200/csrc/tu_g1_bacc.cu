// G1 bucket sums with several lanes per bucket (field arithmetic inlined, as in the chunked accumulation)
#define ZK_NO_FSQR
#include "bucket_acc.cuh"
namespace zk {
cudaError_t zk_bucket_acc_g1(cudaStream_t st, const MsmPlan& p, const uint32_t* offsets, const Entry* entries,
                             const Affine<Fp>* points, uint32_t direct, XYZZ<Fp>* bucket_sums, XYZZ<Fp>* lane_sums,
                             uint32_t* big) {
  return bucket_acc_launch<G1>(st, p, offsets, entries, points, direct, bucket_sums, lane_sums, big);
}
}  // namespace zk
