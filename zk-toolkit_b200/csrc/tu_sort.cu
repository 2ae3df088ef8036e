// recode / histogram / scatter kernels and the block-cooperative exclusive scan (no field arithmetic)
#define ZK_DEFINE_LAUNCH
#define ZK_FMUL_NOINLINE
#include "launch.cuh"
#include "msm.cuh"
ZK_INSTANTIATE_KERNEL(zk::RecodeCount);
ZK_INSTANTIATE_KERNEL(zk::Scatter);
ZK_INSTANTIATE_KERNEL(zk::BucketSizeCheck);

namespace zk {

// 256 threads x 4 consecutive elements (one 128-bit load) = SCAN_SEG elements per block
static constexpr int SCAN_THREADS = 256;
static_assert(SCAN_THREADS * 4 == SCAN_SEG, "scan tile");

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
  __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, inc, d);
    if (lane >= d) inc += t;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < SCAN_THREADS / 32 ? warp_sums[lane] : 0, winc = w;
#pragma unroll
    for (int d = 1; d < SCAN_THREADS / 32; d <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, winc, d);
      if (lane >= d) winc += t;
    }
    if (lane < SCAN_THREADS / 32) warp_sums[lane] = winc - w;  // exclusive warp bases
    if (lane == SCAN_THREADS / 32 - 1) *total = winc;
  }
  __syncthreads();
  return inc - v + warp_sums[warp];
}

// four consecutive scan inputs starting at i.  PAIRS: the input is not stored anywhere -- element j is the number of
// outputs bucket j leaves in a batched-affine round, ceil((off[j + 1] - off[j]) / 2), read from the offsets (n + 1 of
// them) of the round before (this fuses the former PairCount launch into the scan)
template <bool PAIRS> __device__ __forceinline__ uint4 load4(const uint32_t* p, uint32_t i, uint32_t n) {
  uint4 v = make_uint4(0, 0, 0, 0);
  if (PAIRS) {
    uint32_t o[5];
#pragma unroll
    for (int k = 0; k < 5; k++) o[k] = i + k <= n ? p[i + k] : 0;
    if (i < n) v.x = (o[1] - o[0] + 1) / 2;
    if (i + 1 < n) v.y = (o[2] - o[1] + 1) / 2;
    if (i + 2 < n) v.z = (o[3] - o[2] + 1) / 2;
    if (i + 3 < n) v.w = (o[4] - o[3] + 1) / 2;
    return v;
  }
  if (i + 3 < n) return *reinterpret_cast<const uint4*>(p + i);
  if (i < n) v.x = p[i];
  if (i + 1 < n) v.y = p[i + 1];
  if (i + 2 < n) v.z = p[i + 2];
  return v;
}

template <bool PAIRS>
__global__ void __launch_bounds__(SCAN_THREADS) scan_block_sums(uint32_t n, const uint32_t* in, uint32_t* blocksums) {
  __shared__ uint32_t total;
  uint32_t i = (blockIdx.x * SCAN_THREADS + threadIdx.x) * 4;
  uint4 v = load4<PAIRS>(in, i, n);
  block_exclusive_scan(v.x + v.y + v.z + v.w, &total);
  if (threadIdx.x == 0) blocksums[blockIdx.x] = total;
}

// one block: exclusive scan of up to 8 * SCAN_THREADS block sums in place; grand total -> *grand
__global__ void __launch_bounds__(SCAN_THREADS) scan_top_level(uint32_t nblocks, uint32_t* blocksums, uint32_t* grand) {
  __shared__ uint32_t total;
  uint32_t run = 0;
  for (uint32_t base = 0; base < nblocks; base += SCAN_THREADS) {   // tiles in order, carried base
    uint32_t i = base + threadIdx.x;
    uint32_t v = i < nblocks ? blocksums[i] : 0;
    uint32_t ex = block_exclusive_scan(v, &total);
    if (i < nblocks) blocksums[i] = run + ex;
    run += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *grand = run;
}

// cursor (nullable): second copy of the offsets (the scatter cursors of the counting sort)
template <bool PAIRS>
__global__ void __launch_bounds__(SCAN_THREADS) scan_apply(uint32_t n, const uint32_t* in, uint32_t* cursor, const uint32_t* blocksums,
                                                           uint32_t* offsets) {
  __shared__ uint32_t total;
  uint32_t i = (blockIdx.x * SCAN_THREADS + threadIdx.x) * 4;
  uint4 v = load4<PAIRS>(in, i, n);
  uint32_t ex = block_exclusive_scan(v.x + v.y + v.z + v.w, &total) + blocksums[blockIdx.x];
  uint4 o = make_uint4(ex, ex + v.x, ex + v.x + v.y, ex + v.x + v.y + v.z);
  if (i + 3 < n) {
    *reinterpret_cast<uint4*>(offsets + i) = o;
    if (cursor) *reinterpret_cast<uint4*>(cursor + i) = o;
  } else {
    if (i < n) { offsets[i] = o.x; if (cursor) cursor[i] = o.x; }
    if (i + 1 < n) { offsets[i + 1] = o.y; if (cursor) cursor[i + 1] = o.y; }
    if (i + 2 < n) { offsets[i + 2] = o.z; if (cursor) cursor[i + 2] = o.z; }
  }
}

// whole scan in ONE block for n <= kSingleMax (the bucket counts of a per-rank share at 8 GPUs; at 16384 elements the three-kernel form is faster): one launch
// instead of three, tiles of SCAN_SEG elements in order with a carried base
static constexpr uint32_t kSingleMax = 8192;
template <bool PAIRS>
__global__ void __launch_bounds__(SCAN_THREADS) scan_single_block(uint32_t n, const uint32_t* in, uint32_t* cursor, uint32_t* offsets) {
  __shared__ uint32_t total;
  uint32_t run = 0;
  for (uint32_t base = 0; base < n; base += SCAN_SEG) {
    uint32_t i = base + threadIdx.x * 4;
    uint4 v = load4<PAIRS>(in, i, n);
    uint32_t ex = block_exclusive_scan(v.x + v.y + v.z + v.w, &total) + run;
    uint32_t o[4] = {ex, ex + v.x, ex + v.x + v.y, ex + v.x + v.y + v.z};
#pragma unroll
    for (int k = 0; k < 4; k++)
      if (i + k < n) { offsets[i + k] = o[k]; if (cursor) cursor[i + k] = o[k]; }
    run += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) offsets[n] = run;
}

template <bool PAIRS>
static cudaError_t scan_launch(cudaStream_t st, uint32_t n, const uint32_t* in, uint32_t* cursor, uint32_t* offsets, uint32_t* blocksums) {
  if (n <= kSingleMax) {
    scan_single_block<PAIRS><<<1, SCAN_THREADS, 0, st>>>(n, in, cursor, offsets);
    return cudaGetLastError();
  }
  uint32_t nblocks = (n + SCAN_SEG - 1) / SCAN_SEG;
  scan_block_sums<PAIRS><<<nblocks, SCAN_THREADS, 0, st>>>(n, in, blocksums);
  scan_top_level<<<1, SCAN_THREADS, 0, st>>>(nblocks, blocksums, offsets + n);
  scan_apply<PAIRS><<<nblocks, SCAN_THREADS, 0, st>>>(n, in, cursor, blocksums, offsets);
  return cudaGetLastError();
}

// offsets[0..n] = exclusive scan of hist; hist <- offsets (the scatter cursors)
cudaError_t zk_exclusive_scan(cudaStream_t st, uint32_t n, uint32_t* hist_cursor, uint32_t* offsets, uint32_t* blocksums) {
  return scan_launch<false>(st, n, hist_cursor, hist_cursor, offsets, blocksums);
}
// off_out[0..n] = exclusive scan of ceil((off_in[j + 1] - off_in[j]) / 2): the output offsets of a batched-affine round
cudaError_t zk_exclusive_scan_pairs(cudaStream_t st, uint32_t n, const uint32_t* off_in, uint32_t* off_out, uint32_t* blocksums) {
  return scan_launch<true>(st, n, off_in, nullptr, off_out, blocksums);
}
int zk_scan_launches(uint32_t n) { return n <= kSingleMax ? 1 : 3; }

}  // namespace zk
