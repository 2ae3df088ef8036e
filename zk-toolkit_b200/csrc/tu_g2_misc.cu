// G2 kernels off the hot loop: fix-up, reduction, finish, point-set preparation
#define ZK_DEFINE_LAUNCH
#define ZK_FMUL_NOINLINE
#include "launch.cuh"
#include "msm.cuh"
ZK_INSTANTIATE_KERNEL(zk::FixupLevel<zk::G2>);
ZK_INSTANTIATE_KERNEL(zk::BucketReduce<zk::G2>);
ZK_INSTANTIATE_KERNEL(zk::PairSum<zk::G2>);
ZK_INSTANTIATE_KERNEL(zk::Finish<zk::G2>);
ZK_INSTANTIATE_KERNEL(zk::CombinePartials<zk::G2>);
ZK_INSTANTIATE_KERNEL(zk::LoadPoints<zk::G2>);
ZK_INSTANTIATE_KERNEL(zk::StorePoints<zk::G2>);
ZK_INSTANTIATE_KERNEL(zk::PrecomputeSlabs<zk::G2>);
ZK_INSTANTIATE_KERNEL(zk::BaseTableChain<zk::G2>);
ZK_INSTANTIATE_KERNEL(zk::BaseTableAffine<zk::G2>);
ZK_INSTANTIATE_KERNEL(zk::FixedBaseMul<zk::G2>);
