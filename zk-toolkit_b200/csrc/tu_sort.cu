// recode / histogram / scan / scatter kernels (no field arithmetic)
#define ZK_DEFINE_LAUNCH
#define ZK_FMUL_NOINLINE
#include "launch.cuh"
#include "msm.cuh"
ZK_INSTANTIATE_KERNEL(zk::RecodeCount);
ZK_INSTANTIATE_KERNEL(zk::ScanLocal);
ZK_INSTANTIATE_KERNEL(zk::ScanTop);
ZK_INSTANTIATE_KERNEL(zk::ScanApply);
ZK_INSTANTIATE_KERNEL(zk::Scatter);
