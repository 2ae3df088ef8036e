// G2 bucket accumulation (Fq multiplication as a call: 28 per mixed add)
#define ZK_DEFINE_LAUNCH
#define ZK_FMUL_NOINLINE
#include "launch.cuh"
#include "msm.cuh"
ZK_INSTANTIATE_KERNEL(zk::Accumulate<zk::G2>);
