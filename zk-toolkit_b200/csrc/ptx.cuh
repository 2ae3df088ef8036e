// Carry-chain integer primitives for sm_100a, one PTX instruction each.
//
// ptxas fuses every (mad.lo.cc, madc.hi.cc) pair on the same operands into one
// IMAD.WIDE.U32.X with a predicate carry (checked with cuobjdump -sass), so a
// 32x32->64 limb product plus 64-bit accumulate plus carry costs ONE fma-pipe slot.
//
// The carry flag lives in PTX's CC register between consecutive `asm volatile`
// statements; nvcc never emits .cc instructions of its own, and volatile keeps our
// statements in program order.
//
// When the file is compiled WITHOUT nvcc (tests/host_emu only: test infrastructure,
// never shipped) each primitive is emulated in portable C++ with an explicit carry
// variable, so the limb algorithms built on top can be checked on a machine with no GPU.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define ZK_HD __host__ __device__ __forceinline__
#define ZK_D __device__ __forceinline__
#define ZK_CONST static __device__ __constant__ const
#else
#define ZK_HD inline
#define ZK_D inline
#define ZK_CONST static const
#endif

namespace zk {
namespace ptx {

#if defined(__CUDA_ARCH__)

ZK_D uint32_t add_cc(uint32_t a, uint32_t b) { uint32_t d; asm volatile("add.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
ZK_D uint32_t addc_cc(uint32_t a, uint32_t b) { uint32_t d; asm volatile("addc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
ZK_D uint32_t addc(uint32_t a, uint32_t b) { uint32_t d; asm volatile("addc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
ZK_D uint32_t sub_cc(uint32_t a, uint32_t b) { uint32_t d; asm volatile("sub.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
ZK_D uint32_t subc_cc(uint32_t a, uint32_t b) { uint32_t d; asm volatile("subc.cc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
ZK_D uint32_t subc(uint32_t a, uint32_t b) { uint32_t d; asm volatile("subc.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
ZK_D uint32_t mul_lo(uint32_t a, uint32_t b) { uint32_t d; asm volatile("mul.lo.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
ZK_D uint32_t mul_hi(uint32_t a, uint32_t b) { uint32_t d; asm volatile("mul.hi.u32 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b)); return d; }
ZK_D uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("mad.lo.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
ZK_D uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("mad.hi.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
ZK_D uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("madc.lo.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
ZK_D uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("madc.hi.cc.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }
ZK_D uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { uint32_t d; asm volatile("madc.hi.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c)); return d; }

#else  // host emulation of the same instruction semantics (also parsed by nvcc's host pass)

// one flag per thread: carry for add/mad chains, borrow for sub chains (as PTX CC.CF)
inline uint32_t& cc_flag() { static thread_local uint32_t cc = 0; return cc; }

inline uint32_t add_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b; cc_flag() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a + b + cc_flag(); cc_flag() = (uint32_t)(t >> 32); return (uint32_t)t; }
inline uint32_t addc(uint32_t a, uint32_t b) { return a + b + cc_flag(); }
inline uint32_t sub_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b; cc_flag() = (uint32_t)(t >> 63); return (uint32_t)t; }
inline uint32_t subc_cc(uint32_t a, uint32_t b) { uint64_t t = (uint64_t)a - b - cc_flag(); cc_flag() = (uint32_t)(t >> 63); return (uint32_t)t; }
inline uint32_t subc(uint32_t a, uint32_t b) { return a - b - cc_flag(); }
inline uint32_t mul_lo(uint32_t a, uint32_t b) { return a * b; }
inline uint32_t mul_hi(uint32_t a, uint32_t b) { return (uint32_t)(((uint64_t)a * b) >> 32); }
inline uint32_t mad_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(mul_lo(a, b), c); }
inline uint32_t mad_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return add_cc(mul_hi(a, b), c); }
inline uint32_t madc_lo_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(mul_lo(a, b), c); }
inline uint32_t madc_hi_cc(uint32_t a, uint32_t b, uint32_t c) { return addc_cc(mul_hi(a, b), c); }
inline uint32_t madc_hi(uint32_t a, uint32_t b, uint32_t c) { return addc(mul_hi(a, b), c); }

#endif

}  // namespace ptx
}  // namespace zk
