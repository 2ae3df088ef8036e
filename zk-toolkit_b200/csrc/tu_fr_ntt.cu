// Several radix-2 stages of an Fr transform per pass through shared memory (the per-stage bodies of fr_ntt.cuh are
// one global-memory round trip each: 21 of them for a 2^21-point transform; here three).
//
// One pass applies the k stages with block lengths L0, L0/2, ..., L0 >> (k-1) (forward, decimation in frequency;
// the inverse applies the same stages in the opposite order, decimation in time).  Those stages only ever combine
// elements whose indices differ in k fixed bits: a "group" is the 2^k elements  blk L0 + pos + j sub,  j < 2^k,
// sub = L0 >> k.  A block of 256 threads keeps G groups with neighbouring pos (or, when sub = 1, G neighbouring
// contiguous groups) in shared memory, limb-major so that a warp touching consecutive elements is conflict-free;
// global accesses are whole 32-byte sectors, 8 of them contiguous whenever sub >= G.
#include <cuda_runtime.h>
#include "fr_ntt.cuh"

namespace zk {

static constexpr int kNttThreads = 256;

template <bool INV>
__global__ void __launch_bounds__(kNttThreads) fr_ntt_fused_kernel(Fr* a, uint32_t L0, uint32_t k, uint32_t G,
                                                                  uint32_t m_over_L0, const Fr* tw) {
  extern __shared__ __align__(16) uint32_t sm[];
  const uint32_t E = G << k, sub = L0 >> k, gid0 = blockIdx.x * G;
  auto sm_load = [&](uint32_t e) { Fr r;
#pragma unroll
    for (int w = 0; w < 8; w++) r.v[w] = sm[w * E + e];
    return r; };
  auto sm_store = [&](uint32_t e, const Fr& r) {
#pragma unroll
    for (int w = 0; w < 8; w++) sm[w * E + e] = r.v[w]; };
  auto locate = [&](uint32_t idx, uint32_t& e) -> size_t {
    uint32_t g, j;
    if (sub >= G) { g = idx % G; j = idx / G; } else { j = idx & ((1u << k) - 1); g = idx >> k; }
    const uint32_t gid = gid0 + g;
    e = (g << k) + j;
    return (size_t)(gid / sub) * L0 + (gid & (sub - 1)) + (size_t)j * sub;
  };
  for (uint32_t idx = threadIdx.x; idx < E; idx += kNttThreads) {
    uint32_t e;
    size_t addr = locate(idx, e);
    sm_store(e, a[addr]);
  }
  __syncthreads();
  for (uint32_t st = 0; st < k; st++) {
    const uint32_t s = INV ? k - 1 - st : st, bit = k - 1 - s, len = L0 >> s;
    for (uint32_t b = threadIdx.x; b < E / 2; b += kNttThreads) {
      const uint32_t g = b >> (k - 1), jj = b & ((1u << (k - 1)) - 1);
      const uint32_t j_lo = ((jj >> bit) << (bit + 1)) | (jj & ((1u << bit) - 1)), j_hi = j_lo | (1u << bit);
      const uint32_t pos = (gid0 + g) & (sub - 1);
      const uint32_t pin = (pos + j_lo * sub) & (len - 1);           // position inside the block of this stage
      const uint32_t e_lo = (g << k) + j_lo, e_hi = (g << k) + j_hi;
      Fr x = sm_load(e_lo), y = sm_load(e_hi), sum, dif;
      if (INV) {
        if (pin) { Fr w = tw[(size_t)pin * (m_over_L0 << s)]; fmul(y, y, w); }
        fadd(sum, x, y);
        fsub(dif, x, y);
      } else {
        fadd(sum, x, y);
        fsub(dif, x, y);
        if (pin) { Fr w = tw[(size_t)pin * (m_over_L0 << s)]; fmul(dif, dif, w); }
      }
      sm_store(e_lo, sum);
      sm_store(e_hi, dif);
    }
    __syncthreads();
  }
  for (uint32_t idx = threadIdx.x; idx < E; idx += kNttThreads) {
    uint32_t e;
    size_t addr = locate(idx, e);
    a[addr] = sm_load(e);
  }
}

// per-device opt-in (zkmsm_create), see launch.cuh
cudaError_t zk_opt_in_shared_memory_ntt() {
  cudaError_t e = cudaFuncSetAttribute(fr_ntt_fused_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(fr_ntt_fused_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  return e;
}

// returns cudaErrorNotSupported when the shape does not fit (caller falls back to the per-stage bodies)
cudaError_t zk_ntt_fused(cudaStream_t st, bool inverse, Fr* a, uint32_t total, uint32_t L0, uint32_t k, const Fr* tw,
                         uint32_t m) {
  const uint32_t G = 8, sub = L0 >> k, groups = total >> k;
  if (k < 2 || k > 8 || (L0 >> k) << k != L0 || groups % G != 0 || !(sub == 1 || sub % G == 0) || total % L0 != 0)
    return cudaErrorNotSupported;
  const size_t smem = (size_t)(G << k) * sizeof(Fr);
  if (inverse) fr_ntt_fused_kernel<true><<<groups / G, kNttThreads, smem, st>>>(a, L0, k, G, m / L0, tw);
  else fr_ntt_fused_kernel<false><<<groups / G, kNttThreads, smem, st>>>(a, L0, k, G, m / L0, tw);
  return cudaGetLastError();
}

}  // namespace zk
