// G1 batched-affine pre-reduction rounds (field arithmetic inlined; 3 blocks of 128 threads per SM)
#define ZK_DEFINE_LAUNCH
#define ZK_MIN_BLOCKS 3
#include "launch.cuh"
#include "msm.cuh"
ZK_INSTANTIATE_KERNEL(zk::BatchedAddRound<zk::G1, true>);
ZK_INSTANTIATE_KERNEL(zk::BatchedAddRound<zk::G1, false>);
