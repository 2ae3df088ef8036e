// G1 block-cooperative bucket reduction and window tree (coop.cuh)
#include <cuda_runtime.h>
#include "coop.cuh"

namespace zk {

template <class K> static cudaError_t opt_in_smem(K kernel, size_t bytes) {
  static bool done = false;   // one device per process (one process per GPU)
  if (done) return cudaSuccess;
  cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) done = true;
  return e;
}

cudaError_t zk_coop_bucket_reduce_g1(cudaStream_t st, const MsmPlan& p, const uint32_t* offsets, const XYZZ<Fp>* buckets,
                                     XYZZ<Fp>* out) {
  const size_t smem = coop::smem_bytes<Fp>();
  cudaError_t e = opt_in_smem(coop::bucket_reduce_kernel<G1>, smem);
  if (e != cudaSuccess) return e;
  uint32_t chains = p.nwin * (p.B / p.K);
  coop::bucket_reduce_kernel<G1><<<(chains + 31) / 32, coop::kThreads, smem, st>>>(p, offsets, buckets, out);
  return cudaGetLastError();
}

cudaError_t zk_coop_pair_sum_g1(cudaStream_t st, uint32_t nwin, uint32_t pitch, uint32_t m, uint32_t half, XYZZ<Fp>* arr) {
  const size_t smem = coop::smem_bytes<Fp>();
  cudaError_t e = opt_in_smem(coop::pair_sum_kernel<G1>, smem);
  if (e != cudaSuccess) return e;
  uint32_t pairs = nwin * half;
  coop::pair_sum_kernel<G1><<<(pairs + 31) / 32, coop::kThreads, smem, st>>>(nwin, pitch, m, half, arr);
  return cudaGetLastError();
}

cudaError_t zk_coop_finish_g1(cudaStream_t st, uint32_t nwin, uint32_t pitch, uint32_t c, const XYZZ<Fp>* arr,
                              XYZZ<Fp>* out_xyzz, uint32_t* out_affine, uint32_t* out_inf) {
  const size_t smem = coop::smem_bytes<Fp>();
  cudaError_t e = opt_in_smem(coop::finish_kernel<G1>, smem);
  if (e != cudaSuccess) return e;
  coop::finish_kernel<G1><<<1, coop::kThreads, smem, st>>>(nwin, pitch, c, arr, out_xyzz, out_affine, out_inf);
  return cudaGetLastError();
}

cudaError_t zk_coop_combine_g1(cudaStream_t st, uint32_t k, const XYZZ<Fp>* parts, uint32_t* out_affine, uint32_t* out_inf) {
  const size_t smem = coop::smem_bytes<Fp>();
  cudaError_t e = opt_in_smem(coop::combine_kernel<G1>, smem);
  if (e != cudaSuccess) return e;
  coop::combine_kernel<G1><<<1, coop::kThreads, smem, st>>>(k, parts, out_affine, out_inf);
  return cudaGetLastError();
}

}  // namespace zk
