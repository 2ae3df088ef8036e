// G1 point-set preparation and fixed-base kernels (field multiplication as a call)
#define ZK_DEFINE_LAUNCH
#define ZK_FMUL_NOINLINE
#include "launch.cuh"
#include "msm.cuh"
ZK_INSTANTIATE_KERNEL(zk::LoadPoints<zk::G1>);
ZK_INSTANTIATE_KERNEL(zk::StorePoints<zk::G1>);
ZK_INSTANTIATE_KERNEL(zk::PrecomputeSlabs<zk::G1>);
ZK_INSTANTIATE_KERNEL(zk::BaseTableChain<zk::G1>);
ZK_INSTANTIATE_KERNEL(zk::BaseTableAffine<zk::G1>);
ZK_INSTANTIATE_KERNEL(zk::FixedBaseMul<zk::G1>);
ZK_INSTANTIATE_KERNEL(zk::SubgroupCheck<zk::G1>);
