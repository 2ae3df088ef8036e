// the hot kernel: G1 bucket accumulation, field arithmetic fully inlined.
// Measured on B200 (profiles/r1_accumulate_variants.log): capping registers to fit a third 128-thread block
// (168 instead of 188) costs 7 % (spills inside the mixed add), and the dedicated squaring costs 4 % here
// (its extra alu work outweighs the 66 saved IMAD.WIDE), so this unit uses fmul for squares and no cap.
#define ZK_DEFINE_LAUNCH
#define ZK_NO_FSQR
#define ZK_ACC_DOUBLE_BUFFER   // next point held in registers under the current mixed add (208 registers, -2.7 %)
#include "launch.cuh"
#include "msm.cuh"
ZK_INSTANTIATE_KERNEL_STRIDED(zk::Accumulate<zk::G1>);   // launched on a capped grid (gated fallback of AccumulateBuckets)
