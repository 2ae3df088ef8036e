// G1 fix-up tree (field arithmetic inlined)
#define ZK_DEFINE_LAUNCH
#include "launch.cuh"
#include "msm.cuh"
ZK_INSTANTIATE_KERNEL_STRIDED(zk::FixupLevel<zk::G1>);   // launched on a capped grid (gated fallback of AccumulateBuckets)
ZK_INSTANTIATE_KERNEL(zk::FixupDirect<zk::G1>);
