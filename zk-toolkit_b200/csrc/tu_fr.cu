// Fr vector kernels (witness aggregation)
#define ZK_DEFINE_LAUNCH
#include "launch.cuh"
#include "fr_ops.cuh"
#include "fr_ntt.cuh"
ZK_INSTANTIATE_KERNEL(zk::FrToMont);
ZK_INSTANTIATE_KERNEL(zk::FrAggregate);
ZK_INSTANTIATE_KERNEL(zk::FrVecToMont);
ZK_INSTANTIATE_KERNEL(zk::Groth16Scalars);
ZK_INSTANTIATE_KERNEL(zk::Groth16TailB);
ZK_INSTANTIATE_KERNEL(zk::FrPolyMulSub);
ZK_INSTANTIATE_KERNEL(zk::FrTStep);
ZK_INSTANTIATE_KERNEL(zk::FrDivStep);
ZK_INSTANTIATE_KERNEL(zk::FrQuotientOut);
// quotient polynomial by transforms (fr_ntt.cuh)
ZK_INSTANTIATE_KERNEL(zk::FrNttSetup);
ZK_INSTANTIATE_KERNEL(zk::FrPowTable);
ZK_INSTANTIATE_KERNEL(zk::FrNttDif);
ZK_INSTANTIATE_KERNEL(zk::FrNttDit);
ZK_INSTANTIATE_KERNEL(zk::FrScalePow2);
ZK_INSTANTIATE_KERNEL(zk::FrPointMul);
ZK_INSTANTIATE_KERNEL(zk::FrCopyPad);
ZK_INSTANTIATE_KERNEL(zk::FrTreeLeaves);
ZK_INSTANTIATE_KERNEL(zk::FrTreeExpand);
ZK_INSTANTIATE_KERNEL(zk::FrTreeMul);
ZK_INSTANTIATE_KERNEL(zk::FrTreeCombine);
ZK_INSTANTIATE_KERNEL(zk::FrTreeToT);
ZK_INSTANTIATE_KERNEL(zk::FrRevT);
ZK_INSTANTIATE_KERNEL(zk::FrTwoMinus);
ZK_INSTANTIATE_KERNEL(zk::FrSubW);
ZK_INSTANTIATE_KERNEL(zk::FrRevTop);
ZK_INSTANTIATE_KERNEL(zk::FrExtractH);
ZK_INSTANTIATE_KERNEL(zk::FrCheckEqual);
