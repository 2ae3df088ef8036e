// G2 batched-affine pre-reduction rounds (field arithmetic inlined, as in the accumulation kernels; with the Fq
// multiplication as a call every operand travels through local memory: 1.4 KB of stack and 1.95 ns per addition)
#define ZK_DEFINE_LAUNCH
#define ZK_MIN_BLOCKS 2
#include "launch.cuh"
#include "msm.cuh"
ZK_INSTANTIATE_KERNEL(zk::BatchedAddRound<zk::G2, true>);
ZK_INSTANTIATE_KERNEL(zk::BatchedAddRound<zk::G2, false>);
