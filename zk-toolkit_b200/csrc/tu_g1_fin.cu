// G1 Horner / affine conversion / partial combination.  Single-thread kernels that run once per MSM on cold
// instruction caches: the field multiplication is a call (one 6 KB copy that stays hot) instead of ~20 inlined copies.
#define ZK_DEFINE_LAUNCH
#define ZK_FMUL_NOINLINE
#include "launch.cuh"
#include "msm.cuh"
ZK_INSTANTIATE_KERNEL(zk::Finish<zk::G1>);
ZK_INSTANTIATE_KERNEL(zk::CombinePartials<zk::G1>);
ZK_INSTANTIATE_KERNEL(zk::PoisonPartial<zk::G1>);
