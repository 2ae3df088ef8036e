// G1 kernels off the hot loop: fix-up, reduction, finish, point-set preparation
#define ZK_DEFINE_LAUNCH
#define ZK_FMUL_NOINLINE
#include "launch.cuh"
#include "msm.cuh"
ZK_INSTANTIATE_KERNEL(zk::FixupLevel<zk::G1>);
ZK_INSTANTIATE_KERNEL(zk::BucketReduce<zk::G1>);
ZK_INSTANTIATE_KERNEL(zk::PairSum<zk::G1>);
ZK_INSTANTIATE_KERNEL(zk::Finish<zk::G1>);
ZK_INSTANTIATE_KERNEL(zk::CombinePartials<zk::G1>);
ZK_INSTANTIATE_KERNEL(zk::LoadPoints<zk::G1>);
ZK_INSTANTIATE_KERNEL(zk::StorePoints<zk::G1>);
ZK_INSTANTIATE_KERNEL(zk::PrecomputeSlabs<zk::G1>);
ZK_INSTANTIATE_KERNEL(zk::BaseTableChain<zk::G1>);
ZK_INSTANTIATE_KERNEL(zk::BaseTableAffine<zk::G1>);
ZK_INSTANTIATE_KERNEL(zk::FixedBaseMul<zk::G1>);
