// G2 block-cooperative bucket reduction and window tree (coop.cuh)
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

cudaError_t zk_coop_bucket_reduce_g2(cudaStream_t st, const MsmPlan& p, const uint32_t* offsets, const XYZZ<Fp2>* buckets,
                                     XYZZ<Fp2>* out, int tree) {
  const size_t smem = coop::smem_bytes<Fp2>();
  cudaError_t e = opt_in_smem(coop::bucket_reduce_kernel<G2>, smem);
  if (e != cudaSuccess) return e;
  uint32_t chains = p.nwin * (p.B / p.K);
  coop::bucket_reduce_kernel<G2><<<(chains + 31) / 32, coop::kThreads, smem, st>>>(p, offsets, buckets, out, tree);
  return cudaGetLastError();
}

cudaError_t zk_coop_row_sum_g2(cudaStream_t st, uint32_t nwin, uint32_t pitch_in, uint32_t m, uint32_t per_block,
                               const XYZZ<Fp2>* in, uint32_t pitch_out, XYZZ<Fp2>* out) {
  const size_t smem = coop::smem_bytes<Fp2>();
  cudaError_t e = opt_in_smem(coop::row_sum_kernel<G2>, smem);
  if (e != cudaSuccess) return e;
  uint32_t blocks_per_row = (m + per_block - 1) / per_block;
  coop::row_sum_kernel<G2><<<nwin * blocks_per_row, coop::kThreads, smem, st>>>(pitch_in, m, per_block, blocks_per_row, in,
                                                                               pitch_out, out);
  return cudaGetLastError();
}

}  // namespace zk
