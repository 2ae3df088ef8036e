// Development microbenchmarks for the integer pipe of B200 (not part of the product).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench tools/ubench.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../zk-toolkit_b200/csrc/ec.cuh"
using namespace zk;

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); return 1; } } while (0)

// ---- A: IMAD.WIDE chains with real data dependencies (multiplier = low word of the neighbouring
// accumulator), so ptxas cannot hoist or strength-reduce the products.  MODE 0: mad.wide (64-bit
// accumulate), 1: mul.wide + xor fold (IMAD.WIDE with RZ addend), 2: mad.lo.u32 (32-bit IMAD)
template <int NCH, int MODE> __global__ void k_wide_indep(uint32_t* sink, int iters) {
  uint32_t x = threadIdx.x * 2654435761u + blockIdx.x, y = x ^ 0x9e3779b9u;
  uint64_t a[NCH];
#pragma unroll
  for (int k = 0; k < NCH; k++) a[k] = ((uint64_t)(x + k) << 32) | (y + k);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int r = 0; r < 32 / NCH; r++)
#pragma unroll
      for (int k = 0; k < NCH; k++) {
        uint32_t m = (uint32_t)a[(k + 1) % NCH];
        if (MODE == 0) asm("mad.wide.u32 %0, %1, %2, %0;" : "+l"(a[k]) : "r"(m), "r"(y));
        else if (MODE == 1) { uint64_t p; asm("mul.wide.u32 %0, %1, %2;" : "=l"(p) : "r"(m), "r"(y)); a[k] ^= p; }
        else { uint32_t lo = (uint32_t)a[k]; asm("mad.lo.u32 %0, %1, %2, %0;" : "+r"(lo) : "r"(m), "r"(y)); a[k] = (a[k] & 0xffffffff00000000ull) | lo; }
      }
  }
  uint64_t s = 0;
#pragma unroll
  for (int k = 0; k < NCH; k++) s ^= a[k];
  if (s == 0x123456789ull) sink[0] = (uint32_t)s;
}

// ---- B: carry chains (mad.lo.cc / madc.hi.cc pairs -> IMAD.WIDE.U32.X), NCH independent 6-digit chains
template <int NCH> __global__ void k_wide_x(uint32_t* sink, int iters) {
  uint32_t x = threadIdx.x * 2654435761u + blockIdx.x, y = x ^ 0x9e3779b9u;
  uint32_t acc[NCH][12], xs[6];
#pragma unroll
  for (int c = 0; c < NCH; c++)
#pragma unroll
    for (int k = 0; k < 12; k++) acc[c][k] = x + k + c;
#pragma unroll
  for (int k = 0; k < 6; k++) xs[k] = x * (k + 3);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int c = 0; c < NCH; c++) detail::row_mad<12>(acc[c], detail::Arr{xs}, y + c);
  }
  uint32_t s = 0;
#pragma unroll
  for (int c = 0; c < NCH; c++)
#pragma unroll
    for (int k = 0; k < 12; k++) s ^= acc[c][k];
  if (s == 0x12345678u) sink[0] = s;
}

// ---- D: IADD3 carry chains: 12-limb add.cc chains, NCH independent
template <int NCH> __global__ void k_addc(uint32_t* sink, int iters) {
  uint32_t x = threadIdx.x * 2654435761u + blockIdx.x;
  uint32_t acc[NCH][12], b[12];
#pragma unroll
  for (int k = 0; k < 12; k++) b[k] = x * (k + 7);
#pragma unroll
  for (int c = 0; c < NCH; c++)
#pragma unroll
    for (int k = 0; k < 12; k++) acc[c][k] = x + k + c;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int c = 0; c < NCH; c++) {
      acc[c][0] = ptx::add_cc(acc[c][0], b[0]);
#pragma unroll
      for (int k = 1; k < 11; k++) acc[c][k] = ptx::addc_cc(acc[c][k], b[k]);
      acc[c][11] = ptx::addc(acc[c][11], b[11]);
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int c = 0; c < NCH; c++)
#pragma unroll
    for (int k = 0; k < 12; k++) s ^= acc[c][k];
  if (s == 0x12345678u) sink[0] = s;
}

// ---- F: plain IMAD.WIDE product + 64-bit add with carry chain on the alu pipe
__global__ void k_wide_plus_iadd(uint32_t* sink, int iters) {
  uint32_t x = threadIdx.x * 2654435761u + blockIdx.x, y = x ^ 0x9e3779b9u;
  uint32_t acc[2][12], xs[6];
#pragma unroll
  for (int c = 0; c < 2; c++)
#pragma unroll
    for (int k = 0; k < 12; k++) acc[c][k] = x + k + c;
#pragma unroll
  for (int k = 0; k < 6; k++) xs[k] = x * (k + 3);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int c = 0; c < 2; c++) {
      uint64_t p[6];
#pragma unroll
      for (int k = 0; k < 6; k++) asm volatile("mul.wide.u32 %0, %1, %2;" : "=l"(p[k]) : "r"(xs[k]), "r"(y + c));
      acc[c][0] = ptx::add_cc(acc[c][0], (uint32_t)p[0]);
      acc[c][1] = ptx::addc_cc(acc[c][1], (uint32_t)(p[0] >> 32));
#pragma unroll
      for (int k = 1; k < 6; k++) {
        acc[c][2 * k] = ptx::addc_cc(acc[c][2 * k], (uint32_t)p[k]);
        acc[c][2 * k + 1] = ptx::addc_cc(acc[c][2 * k + 1], (uint32_t)(p[k] >> 32));
      }
      y += ptx::addc(0, 0);
    }
  }
  uint32_t s = 0;
#pragma unroll
  for (int c = 0; c < 2; c++)
#pragma unroll
    for (int k = 0; k < 12; k++) s ^= acc[c][k];
  if (s == 0x12345678u) sink[0] = s;
}

// ---- E: modular multiplication throughput: x = x * y repeated
__global__ void k_fmul32(Fp* io, int iters) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  Fp x = io[i], y = io[i ^ 1];
  for (int it = 0; it < iters; it++) fmul(x, x, y);
  io[i] = x;
}

// prototype: 14 limbs of 28 bits, carry-free column accumulation with plain IMAD.WIDE.U32
static __device__ __constant__ uint32_t Q28[14] = {
  0xfffaaab, 0xfefffff, 0x3ffffb9, 0xfffeb15, 0x6241eab, 0xa0f6b0f, 0xf6730d2, 0xf38512b, 0x4774b84, 0x4bacd76, 0xba7b643, 0xe69a4b1, 0x1ea397f, 0x1a011};
// filled by host: q in radix 2^28
#define MASK28 0x0fffffffu
__device__ __forceinline__ uint64_t madw(uint32_t a, uint32_t b, uint64_t c) {
  uint64_t d; asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c)); return d;
}
struct F28 { uint32_t v[14]; };
__device__ __forceinline__ void fmul28(F28& r, const F28& a, const F28& b, uint32_t inv28, const uint32_t* q) {
  uint64_t t[28];
#pragma unroll
  for (int k = 0; k < 28; k++) t[k] = 0;
#pragma unroll
  for (int i = 0; i < 14; i++) {
#pragma unroll
    for (int j = 0; j < 14; j++) t[i + j] = madw(a.v[j], b.v[i], t[i + j]);
    uint32_t m = ((uint32_t)t[i] * inv28) & MASK28;
#pragma unroll
    for (int j = 0; j < 14; j++) t[i + j] = madw(m, q[j], t[i + j]);
    t[i + 1] += t[i] >> 28;
  }
  uint64_t c = 0;
#pragma unroll
  for (int k = 0; k < 14; k++) { c += t[14 + k]; r.v[k] = (uint32_t)c & MASK28; c >>= 28; }
}
__global__ void k_fmul28(F28* io, int iters, uint32_t inv28) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  F28 x = io[i], y = io[i ^ 1];
  uint32_t q[14];
#pragma unroll
  for (int k = 0; k < 14; k++) q[k] = Q28[k];
  for (int it = 0; it < iters; it++) fmul28(x, x, y, inv28, q);
  io[i] = x;
}
// same with q as compile-time immediates
__global__ void k_fmul28c(F28* io, int iters, uint32_t inv28) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  F28 x = io[i], y = io[i ^ 1];
  for (int it = 0; it < iters; it++) fmul28(x, x, y, inv28, Q28);
  io[i] = x;
}

// ---- G: latency of one point operation on a lone warp (what bounds the reduction tail)
template <int MODE> __global__ void k_pointop(XYZZ<Fp>* io, int iters) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  XYZZ<Fp> a = io[2 * i], b = io[2 * i + 1];
  for (int it = 0; it < iters; it++) {
    if (MODE == 0) xyzz_add(a, b);
    else if (MODE == 1) xyzz_add_ilp(a, b);
    else if (MODE == 2) xyzz_dbl(a);
    else xyzz_dbl_ilp(a);
  }
  io[2 * i] = a;
}
__global__ void k_fmul_lat(Fp* io, int iters) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  Fp x = io[i], y = io[i ^ 1];
  for (int it = 0; it < iters; it++) fmul(x, x, y);
  io[i] = x;
}
__global__ void k_fmul2_lat(Fp* io, int iters) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  Fp x = io[i], y = io[i ^ 1], z = io[i ^ 2];
  for (int it = 0; it < iters; it++) fmul2(x, x, y, z, z, y);
  io[i] = x; io[i ^ 2] = z;
}

// MODE 0: finv (batched division steps), 1: finv_euclid, 2: xyzz_to_affine + store_canonical; one thread
template <int MODE> __global__ void k_inv(Fp* io, uint32_t* out, int iters) {
  Fp x = io[0], y;
  if (MODE == 2) {
    XYZZ<Fp> p; p.x = io[0]; p.y = io[1]; p.zz = io[2]; p.zzz = io[3];
    for (int it = 0; it < iters; it++) { Affine<Fp> a; xyzz_to_affine(a, p); p.x = a.x; p.y = a.y; }
    io[4] = p.x; io[5] = p.y;
    return;
  }
  for (int it = 0; it < iters; it++) { if (MODE == 0) finv(y, x); else finv_euclid(y, x); fadd(x, y, io[1]); }
  io[4] = x;
}

template <class K, class... A> float timeit(int grid, int block, K k, A... a) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<<<grid, block>>>(a...);  // warm
  cudaEventRecord(e0);
  k<<<grid, block>>>(a...);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms;
}

int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  uint32_t* sink; CK(cudaMalloc(&sink, 1 << 20));
  int iters = 4096;
  for (int wpb : {128, 256, 512}) {
    int grid = sms * (2048 / wpb);
    double thr = (double)grid * wpb;
    printf("--- block %d, grid %d (full occupancy target)\n", wpb, grid);
    float ms;
    ms = timeit(grid, wpb, k_wide_indep<8, 0>, sink, iters); printf("A mad.wide x8        : %7.2f T/s\n", thr * iters * 32 / ms / 1e9);
    ms = timeit(grid, wpb, k_wide_indep<8, 1>, sink, iters); printf("A mul.wide+xor x8    : %7.2f T/s\n", thr * iters * 32 / ms / 1e9);
    ms = timeit(grid, wpb, k_wide_indep<8, 2>, sink, iters); printf("A mad.lo x8          : %7.2f T/s\n", thr * iters * 32 / ms / 1e9);
    ms = timeit(grid, wpb, k_wide_indep<4, 0>, sink, iters); printf("A mad.wide x4        : %7.2f T/s\n", thr * iters * 32 / ms / 1e9);
    ms = timeit(grid, wpb, k_wide_x<1>, sink, iters); printf("B wide.X 1 chain     : %7.2f T/s\n", thr * iters * 6 / ms / 1e9);
    ms = timeit(grid, wpb, k_wide_x<2>, sink, iters); printf("B wide.X 2 chains    : %7.2f T/s\n", thr * iters * 12 / ms / 1e9);
    ms = timeit(grid, wpb, k_wide_x<4>, sink, iters); printf("B wide.X 4 chains    : %7.2f T/s\n", thr * iters * 24 / ms / 1e9);
    ms = timeit(grid, wpb, k_addc<2>, sink, iters); printf("D addc 2 chains      : %7.2f T/s (adds)\n", thr * iters * 24 / ms / 1e9);
    ms = timeit(grid, wpb, k_addc<4>, sink, iters); printf("D addc 4 chains      : %7.2f T/s (adds)\n", thr * iters * 48 / ms / 1e9);
    ms = timeit(grid, wpb, k_wide_plus_iadd, sink, iters); printf("F mul.wide+addc      : %7.2f T/s (products)\n", thr * iters * 12 / ms / 1e9);
  }
  // modmul throughput at several occupancies
  uint32_t inv28 = 0; { uint32_t q0 = 0xfffaaab; uint32_t x = 1; for (int i = 0; i < 6; i++) x *= 2 - q0 * x; inv28 = (0u - x) & MASK28; }
  Fp* io; CK(cudaMalloc(&io, sizeof(Fp) * sms * 2048)); CK(cudaMemset(io, 0x5a, sizeof(Fp) * sms * 2048));
  F28* io28; CK(cudaMalloc(&io28, sizeof(F28) * sms * 2048)); CK(cudaMemset(io28, 0x05, sizeof(F28) * sms * 2048));
  for (int tps : {128, 256, 384, 512, 768, 1024}) {   // threads per SM
    int block = 128, grid = sms * tps / block, it2 = 2000;
    float ms = timeit(grid, block, k_fmul32, io, it2);
    float ms2 = timeit(grid, block, k_fmul28, io28, it2, inv28);
    float ms3 = timeit(grid, block, k_fmul28c, io28, it2, inv28);
    double n = (double)grid * block * it2;
    printf("modmul @%4d thr/SM: 12x32 carry-chain %7.2f G/s | 14x28 carry-free %7.2f G/s | 14x28 const-q %7.2f G/s\n", tps, n / ms / 1e6, n / ms2 / 1e6, n / ms3 / 1e6);
  }
  {
    XYZZ<Fp>* pio; CK(cudaMalloc(&pio, sizeof(XYZZ<Fp>) * 2 * sms * 128)); CK(cudaMemset(pio, 0x11, sizeof(XYZZ<Fp>) * 2 * sms * 128));
    int it3 = 200;
    for (int warps : {1, 4}) {   // warps per SM (1 = a lone warp on one scheduler, 4 = one per scheduler)
      int block = warps * 32, grid = sms;
      float m0 = timeit(grid, block, k_pointop<0>, pio, it3), m1 = timeit(grid, block, k_pointop<1>, pio, it3);
      float m2 = timeit(grid, block, k_pointop<2>, pio, it3), m3 = timeit(grid, block, k_pointop<3>, pio, it3);
      float f1 = timeit(grid, block, k_fmul_lat, io, 2000), f2 = timeit(grid, block, k_fmul2_lat, io, 2000);
      printf("latency @%d warp(s)/SM: fmul %.3f us | fmul2 pair %.3f us | xyzz_add %.2f us, _ilp %.2f us | xyzz_dbl %.2f us, _ilp %.2f us\n",
             warps, f1 * 1e3 / 2000, f2 * 1e3 / 2000, m0 * 1e3 / it3, m1 * 1e3 / it3, m2 * 1e3 / it3, m3 * 1e3 / it3);
    }
  }
  {
    int it4 = 20;
    float a = timeit(1, 1, k_inv<0>, io, sink, it4), b = timeit(1, 1, k_inv<1>, io, sink, it4), c = timeit(1, 1, k_inv<2>, io, sink, it4);
    printf("one thread: finv (division steps) %.1f us | finv_euclid %.1f us | xyzz_to_affine %.1f us\n", a * 1e3 / it4, b * 1e3 / it4, c * 1e3 / it4);
  }
  CK(cudaDeviceSynchronize());
  return 0;
}
