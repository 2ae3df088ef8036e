// G2 bucket accumulation (Fq multiplication inlined: 28 per mixed add)
#define ZK_DEFINE_LAUNCH

#include "launch.cuh"
#include "msm.cuh"
ZK_INSTANTIATE_KERNEL_STRIDED(zk::Accumulate<zk::G2>);   // launched on a capped grid (gated fallback of AccumulateBuckets)
