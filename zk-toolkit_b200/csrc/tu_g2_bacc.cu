// G2 bucket sums with several lanes per bucket
#include "bucket_acc.cuh"
namespace zk {
cudaError_t zk_bucket_acc_g2(cudaStream_t st, const MsmPlan& p, const uint32_t* offsets, const Entry* entries,
                             const Affine<Fp2>* points, uint32_t direct, XYZZ<Fp2>* bucket_sums, XYZZ<Fp2>* lane_sums,
                             uint32_t* big) {
  return bucket_acc_launch<G2>(st, p, offsets, entries, points, direct, bucket_sums, lane_sums, big);
}
}  // namespace zk
