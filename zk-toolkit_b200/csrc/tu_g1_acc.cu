// the hot kernel: G1 bucket accumulation, field arithmetic fully inlined
#define ZK_DEFINE_LAUNCH
#define ZK_MIN_BLOCKS 3   // 168 registers: three 128-thread blocks per SM (dedicated squaring would otherwise take 188)
#include "launch.cuh"
#include "msm.cuh"
ZK_INSTANTIATE_KERNEL(zk::Accumulate<zk::G1>);
