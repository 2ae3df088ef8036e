// Fr vector kernels (witness aggregation)
#define ZK_DEFINE_LAUNCH
#include "launch.cuh"
#include "fr_ops.cuh"
ZK_INSTANTIATE_KERNEL(zk::FrToMont);
ZK_INSTANTIATE_KERNEL(zk::FrAggregate);
