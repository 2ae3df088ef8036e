// G1 Horner / affine conversion / partial combination (field arithmetic inlined)
#define ZK_DEFINE_LAUNCH
#include "launch.cuh"
#include "msm.cuh"
ZK_INSTANTIATE_KERNEL(zk::Finish<zk::G1>);
ZK_INSTANTIATE_KERNEL(zk::CombinePartials<zk::G1>);
