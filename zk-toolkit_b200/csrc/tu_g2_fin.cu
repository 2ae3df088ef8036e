// G2 Horner / affine conversion / partial combination: single-thread finishing kernels
#define ZK_DEFINE_LAUNCH
#define ZK_FMUL_NOINLINE   // fallback / single-thread kernels: field multiplication as a call keeps the build short
#include "launch.cuh"
#include "msm.cuh"
ZK_INSTANTIATE_KERNEL(zk::Finish<zk::G2>);
ZK_INSTANTIATE_KERNEL(zk::CombinePartials<zk::G2>);
ZK_INSTANTIATE_KERNEL(zk::PoisonPartial<zk::G2>);
