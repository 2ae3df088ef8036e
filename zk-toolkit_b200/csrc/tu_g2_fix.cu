// G2 fix-up (field arithmetic inlined; the call-based multiplication measured slower here: 1.59 against 1.14 ms for
// FixupDirect at 2^18)
#define ZK_DEFINE_LAUNCH
#include "launch.cuh"
#include "msm.cuh"
ZK_INSTANTIATE_KERNEL_STRIDED(zk::FixupLevel<zk::G2>);   // launched on a capped grid (gated fallback of AccumulateBuckets)
ZK_INSTANTIATE_KERNEL(zk::FixupDirect<zk::G2>);
