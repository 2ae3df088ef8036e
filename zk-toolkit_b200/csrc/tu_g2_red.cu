// G2 bucket reduction and pairwise tree (per-thread fallbacks; the cooperative kernels of coop.cuh do the work)
#define ZK_DEFINE_LAUNCH
#define ZK_FMUL_NOINLINE   // fallback / single-thread kernels: field multiplication as a call keeps the build short
#include "launch.cuh"
#include "msm.cuh"
ZK_INSTANTIATE_KERNEL(zk::BucketReduce<zk::G2>);
ZK_INSTANTIATE_KERNEL(zk::PairSum<zk::G2>);
ZK_INSTANTIATE_KERNEL(zk::BucketLaneSum<zk::G2>);
