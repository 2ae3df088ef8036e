// the hot kernel: G1 bucket accumulation, field arithmetic fully inlined
#define ZK_DEFINE_LAUNCH
#include "launch.cuh"
#include "msm.cuh"
ZK_INSTANTIATE_KERNEL(zk::Accumulate<zk::G1>);
