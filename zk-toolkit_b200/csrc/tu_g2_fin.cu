// G2 Horner / affine conversion / partial combination (field arithmetic inlined)
#define ZK_DEFINE_LAUNCH
#include "launch.cuh"
#include "msm.cuh"
ZK_INSTANTIATE_KERNEL(zk::Finish<zk::G2>);
ZK_INSTANTIATE_KERNEL(zk::CombinePartials<zk::G2>);
