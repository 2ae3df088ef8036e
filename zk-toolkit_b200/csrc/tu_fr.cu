// Fr vector kernels (witness aggregation)
#define ZK_DEFINE_LAUNCH
#include "launch.cuh"
#include "fr_ops.cuh"
ZK_INSTANTIATE_KERNEL(zk::FrToMont);
ZK_INSTANTIATE_KERNEL(zk::FrAggregate);
ZK_INSTANTIATE_KERNEL(zk::FrVecToMont);
ZK_INSTANTIATE_KERNEL(zk::FrPolyMulSub);
ZK_INSTANTIATE_KERNEL(zk::FrTStep);
ZK_INSTANTIATE_KERNEL(zk::FrDivStep);
ZK_INSTANTIATE_KERNEL(zk::FrQuotientOut);
