// G2 point-set preparation and fixed-base kernels (field multiplication as a call)
#define ZK_DEFINE_LAUNCH
#define ZK_FMUL_NOINLINE
#include "launch.cuh"
#include "msm.cuh"
ZK_INSTANTIATE_KERNEL(zk::LoadPoints<zk::G2>);
ZK_INSTANTIATE_KERNEL(zk::StorePoints<zk::G2>);
ZK_INSTANTIATE_KERNEL(zk::PrecomputeSlabs<zk::G2>);
ZK_INSTANTIATE_KERNEL(zk::BaseTableChain<zk::G2>);
ZK_INSTANTIATE_KERNEL(zk::BaseTableAffine<zk::G2>);
ZK_INSTANTIATE_KERNEL(zk::FixedBaseMul<zk::G2>);
ZK_INSTANTIATE_KERNEL(zk::SubgroupCheck<zk::G2>);
