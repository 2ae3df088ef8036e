// G2 bucket reduction and pairwise tree (field arithmetic inlined)
#define ZK_DEFINE_LAUNCH
#include "launch.cuh"
#include "msm.cuh"
ZK_INSTANTIATE_KERNEL(zk::BucketReduce<zk::G2>);
ZK_INSTANTIATE_KERNEL(zk::PairSum<zk::G2>);
