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
                                     XYZZ<Fp>* out, int tree) {
  const size_t smem = coop::smem_bytes<Fp>();
  cudaError_t e = opt_in_smem(coop::bucket_reduce_kernel<G1>, smem);
  if (e != cudaSuccess) return e;
  uint32_t chains = p.nwin * (p.B / p.K);
  coop::bucket_reduce_kernel<G1><<<(chains + 31) / 32, coop::kThreads, smem, st>>>(p, offsets, buckets, out, tree);
  return cudaGetLastError();
}

cudaError_t zk_coop_row_sum_g1(cudaStream_t st, uint32_t nwin, uint32_t pitch_in, uint32_t m, uint32_t per_block,
                               const XYZZ<Fp>* in, uint32_t pitch_out, XYZZ<Fp>* out) {
  const size_t smem = coop::smem_bytes<Fp>();
  cudaError_t e = opt_in_smem(coop::row_sum_kernel<G1>, smem);
  if (e != cudaSuccess) return e;
  uint32_t blocks_per_row = (m + per_block - 1) / per_block;
  coop::row_sum_kernel<G1><<<nwin * blocks_per_row, coop::kThreads, smem, st>>>(pitch_in, m, per_block, blocks_per_row, in,
                                                                               pitch_out, out);
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
