// G1 batched-affine pre-reduction rounds (field arithmetic inlined)
#define ZK_DEFINE_LAUNCH
#include "launch.cuh"
#include "msm.cuh"
ZK_INSTANTIATE_KERNEL(zk::PairCount);
ZK_INSTANTIATE_KERNEL(zk::BatchedAddRound<zk::G1, true>);
ZK_INSTANTIATE_KERNEL(zk::BatchedAddRound<zk::G1, false>);
