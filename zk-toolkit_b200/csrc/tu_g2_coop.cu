// G2 block-cooperative bucket reduction and window tree (coop.cuh)
#include <cuda_runtime.h>
#include "coop.cuh"

namespace zk {

// The opt-in to > 48 KB of dynamic shared memory is a per-DEVICE function attribute: zkmsm_create runs this for its
// device (a process may hold contexts on several devices), never the launch path.
cudaError_t zk_opt_in_shared_memory_coop_g2() {
  const size_t bytes = coop::smem_bytes<Fp2>();
  cudaError_t e = cudaSuccess;
  if (e == cudaSuccess) e = cudaFuncSetAttribute(coop::bucket_reduce_kernel<G2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(coop::row_sum_kernel<G2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(coop::combine_kernel<G2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  return e;
}

cudaError_t zk_coop_bucket_reduce_g2(cudaStream_t st, const MsmPlan& p, const uint32_t* offsets, const XYZZ<Fp2>* buckets,
                                     XYZZ<Fp2>* out, int tree) {
  const size_t smem = coop::smem_bytes<Fp2>();
  uint32_t chains = p.nwin * (p.B / p.K);
  coop::bucket_reduce_kernel<G2><<<(chains + 31) / 32, coop::block_threads<Fp2>(), smem, st>>>(p, offsets, buckets, out, tree);
  return cudaGetLastError();
}

cudaError_t zk_coop_row_sum_g2(cudaStream_t st, uint32_t nwin, uint32_t pitch_in, uint32_t m, uint32_t per_block,
                               const XYZZ<Fp2>* in, uint32_t pitch_out, XYZZ<Fp2>* out) {
  const size_t smem = coop::smem_bytes<Fp2>();
  uint32_t blocks_per_row = (m + per_block - 1) / per_block;
  coop::row_sum_kernel<G2><<<nwin * blocks_per_row, coop::block_threads<Fp2>(), smem, st>>>(pitch_in, m, per_block, blocks_per_row, in,
                                                                               pitch_out, out);
  return cudaGetLastError();
}

cudaError_t zk_coop_combine_g2(cudaStream_t st, uint32_t k, const XYZZ<Fp2>* parts, uint32_t stride_words, uint32_t* out_affine,
                               uint32_t* out_inf, uint32_t* err) {
  const size_t smem = coop::smem_bytes<Fp2>();
  coop::combine_kernel<G2><<<1, coop::block_threads<Fp2>(), smem, st>>>(k, parts, stride_words, out_affine, out_inf, err);
  return cudaGetLastError();
}

}  // namespace zk
