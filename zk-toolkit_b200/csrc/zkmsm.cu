// C ABI (include/zkmsm.h) over the sm_100a kernels of msm.cuh.  No CPU compute path exists in
// this library: without a usable CUDA device every entry point returns an error.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <atomic>
#include <new>

#include "../../include/zkmsm.h"
#include "launch.cuh"
#include "msm.cuh"
#include "fr_ops.cuh"
#include "fr_ntt.cuh"

using namespace zk;

// ------------------------------------------------------------------------------------------------

// ------------------------------------------------------------------------------------------------
enum { WS_HIST, WS_OFFSETS, WS_SEGSUM, WS_ENTRIES, WS_BUCKETS, WS_PARTIALS, WS_PKEYS, WS_REDUCED, WS_SCALARS, WS_MISC,
       WS_PRE_A, WS_PRE_B, WS_PRE_PREFIX, WS_PRE_OFF, WS_PRE_ENTRIES, WS_COUNT };

struct alignas(16) ResultBlock {      // device + pinned host mirror
  uint32_t xyzz[96];      // first: XYZZ<F> needs 16-byte alignment
  uint32_t affine[48];
  uint32_t inf;
  uint32_t err;
  uint32_t big;           // AccumulateBuckets' "a bucket is over the cap" flag (device side only)
  uint32_t aux_err;       // zkmsm_groth16_prove: r or s out of range
  uint32_t proof[96];     // zkmsm_groth16_combine: A | B | C
  uint32_t proof_inf[4];
};

// one captured MSM launch sequence (CUDA graph): replayed while the same point set, sizes, buffers and options recur
struct GraphKey {
  uint64_t ps_uid, ws_epoch, tune_epoch;
  size_t n;
  const void* scalars;
  const void* out;
  int mode;             // curve | want_affine << 4
  uint32_t rank, world;
  bool operator==(const GraphKey& o) const {
    return ps_uid == o.ps_uid && ws_epoch == o.ws_epoch && tune_epoch == o.tune_epoch && n == o.n && scalars == o.scalars &&
           out == o.out && mode == o.mode && rank == o.rank && world == o.world;
  }
};
struct GraphSlot { GraphKey key; cudaGraphExec_t exec; int launches; uint64_t used; };
static constexpr int kGraphSlots = 8;

struct zkmsm_ctx {
  int device;
  int sms;
  cudaStream_t own_stream, stream;
  char err[512];
  unsigned window_override;
  MsmTuning tune;       // environment switches, read once in zkmsm_create; zkmsm_set_option overrides
  uint64_t tune_epoch, ws_epoch, graph_clock;
  GraphSlot graphs[kGraphSlots];
  void* ws[WS_COUNT];
  size_t ws_bytes[WS_COUNT];
  ResultBlock* d_res;
  ResultBlock* h_res;
  int last_launches;
  int pending;          // 0 none, 1 kernels enqueued, 2 trivially infinity (n == 0)
  int pending_words;    // 24 or 48
  LaunchProfile* prof;  // non-null while profiling is enabled
  Fr* ntt_tables;       // quotient polynomial: twiddles, transforms of t and of rev(t)^-1 for ntt_n (fr_ntt.cuh)
  size_t ntt_n;
};

static std::atomic<uint64_t> g_next_ps_uid{1};

struct zkmsm_points {
  uint64_t uid;         // never reused (keys the cached launch graphs)
  int curve;            // 1 = G1, 2 = G2
  int device;
  size_t n;             // points per slab
  unsigned c, W;        // window layout of the slabs (precomputed sets)
  bool precomp;
  bool half;            // ZKMSM_SUBGROUP: points have order r, scalars are folded to (r-1)/2
  void* d_pts;          // Affine<F>[W * n] (precomp) or Affine<F>[n]
};

static int fail(zkmsm_ctx* ctx, int code, const char* fmt, ...) {
  if (ctx) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
    va_end(ap);
  }
  return code;
}
#define CU(ctx, call)                                                                                   \
  do {                                                                                                  \
    cudaError_t e__ = (call);                                                                           \
    if (e__ != cudaSuccess) return fail(ctx, ZKMSM_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e__)); \
  } while (0)

static void graphs_drop(zkmsm_ctx* ctx, uint64_t ps_uid /* 0 = all */) {
  for (int i = 0; i < kGraphSlots; i++)
    if (ctx->graphs[i].exec && (ps_uid == 0 || ctx->graphs[i].key.ps_uid == ps_uid)) {
      cudaGraphExecDestroy(ctx->graphs[i].exec);
      ctx->graphs[i].exec = nullptr;
    }
}

static int ws_reserve(zkmsm_ctx* ctx, int slot, size_t bytes) {
  if (bytes == 0) bytes = 16;
  if (ctx->ws_bytes[slot] >= bytes) return ZKMSM_OK;
  ctx->ws_epoch++;          // captured graphs hold the old pointers
  graphs_drop(ctx, 0);
  if (ctx->ws[slot]) {
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    CU(ctx, cudaFree(ctx->ws[slot]));
    ctx->ws[slot] = nullptr;
    ctx->ws_bytes[slot] = 0;
  }
  size_t want = bytes + bytes / 8;
  cudaError_t e = cudaMalloc(&ctx->ws[slot], want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    want = bytes;
    e = cudaMalloc(&ctx->ws[slot], want);
  }
  if (e != cudaSuccess) return fail(ctx, ZKMSM_ERR_NOMEM, "cudaMalloc(%zu): %s", want, cudaGetErrorString(e));
  ctx->ws_bytes[slot] = want;
  return ZKMSM_OK;
}

extern "C" const char* zkmsm_version(void) { return "zkmsm 0.2 (sm_100a; BLS12-381 G1/G2 Pippenger, batched-affine + XYZZ, 12x32-bit Montgomery; Groth16 prove)"; }

extern "C" int zkmsm_create(int device, zkmsm_ctx** out) {
  if (!out) return ZKMSM_ERR_INVALID_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) { cudaGetLastError(); return ZKMSM_ERR_NO_DEVICE; }
  if (device < 0 || device >= count) return ZKMSM_ERR_INVALID_ARG;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return ZKMSM_ERR_CUDA;
  if (prop.major != 10) return ZKMSM_ERR_NO_DEVICE;  // the only code in this library is sm_100a SASS
  zkmsm_ctx* ctx = new (std::nothrow) zkmsm_ctx();
  if (!ctx) return ZKMSM_ERR_NOMEM;
  memset(ctx, 0, sizeof(*ctx));
  ctx->device = device;
  ctx->sms = prop.multiProcessorCount;
  strcpy(ctx->err, "ok");
  if (cudaSetDevice(device) != cudaSuccess || cudaStreamCreateWithFlags(&ctx->own_stream, cudaStreamNonBlocking) != cudaSuccess ||
      cudaMalloc(&ctx->d_res, sizeof(ResultBlock)) != cudaSuccess ||
      cudaMallocHost(&ctx->h_res, sizeof(ResultBlock)) != cudaSuccess) {
    delete ctx;
    return ZKMSM_ERR_CUDA;
  }
  ctx->stream = ctx->own_stream;
  ctx->tune = MsmTuning::from_env();
  // > 48 KB of dynamic shared memory is a per-device opt-in: done here for this context's device
  if (zk_opt_in_shared_memory_coop_g1() != cudaSuccess || zk_opt_in_shared_memory_coop_g2() != cudaSuccess ||
      zk_opt_in_shared_memory_ntt() != cudaSuccess) {
    cudaGetLastError();
    zkmsm_destroy(ctx);
    return ZKMSM_ERR_CUDA;
  }
  *out = ctx;
  return ZKMSM_OK;
}

extern "C" int zkmsm_destroy(zkmsm_ctx* ctx) {
  if (!ctx) return ZKMSM_ERR_INVALID_ARG;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  graphs_drop(ctx, 0);
  for (int i = 0; i < WS_COUNT; i++)
    if (ctx->ws[i]) cudaFree(ctx->ws[i]);
  if (ctx->ntt_tables) cudaFree(ctx->ntt_tables);
  if (ctx->prof) zkmsm_profile(ctx, 0);
  cudaFree(ctx->d_res);
  cudaFreeHost(ctx->h_res);
  cudaStreamDestroy(ctx->own_stream);
  delete ctx;
  return ZKMSM_OK;
}

extern "C" int zkmsm_set_stream(zkmsm_ctx* ctx, void* s) {
  if (!ctx) return ZKMSM_ERR_INVALID_ARG;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  ctx->stream = s ? (cudaStream_t)s : ctx->own_stream;
  return ZKMSM_OK;
}

extern "C" const char* zkmsm_last_error(const zkmsm_ctx* ctx) { return ctx ? ctx->err : "null context"; }

extern "C" int zkmsm_set_window(zkmsm_ctx* ctx, unsigned c) {
  if (!ctx || (c != 0 && (c < 3 || c > 22))) return ZKMSM_ERR_INVALID_ARG;
  ctx->window_override = c;
  ctx->tune_epoch++;   // plain sets plan with it: cached launch graphs are keyed on the epoch
  return ZKMSM_OK;
}

extern "C" int zkmsm_set_option(zkmsm_ctx* ctx, const char* name, long value) {
  if (!ctx || !name) return ZKMSM_ERR_INVALID_ARG;
  struct { const char* name; int MsmTuning::*field; } table[] = {
      {"batch_rounds", &MsmTuning::batch_rounds}, {"batch_T", &MsmTuning::batch_T}, {"batch_g2", &MsmTuning::batch_g2},
      {"batch_blocks", &MsmTuning::batch_blocks}, {"L", &MsmTuning::L}, {"K", &MsmTuning::K},
      {"no_wave_L", &MsmTuning::no_wave_L}, {"no_coop", &MsmTuning::no_coop}, {"ntt_no_fuse", &MsmTuning::ntt_no_fuse},
      {"quotient_schoolbook", &MsmTuning::quotient_schoolbook}, {"no_graph", &MsmTuning::no_graph},
      {"no_bucket_acc", &MsmTuning::no_bucket_acc}, {"acc_G", &MsmTuning::acc_G},
      {"coop_max_chains", &MsmTuning::coop_max_chains}, {"min_left", &MsmTuning::min_left}};
  for (auto& t : table)
    if (strcmp(t.name, name) == 0) {
      ctx->tune.*(t.field) = (int)value;
      ctx->tune_epoch++;
      return ZKMSM_OK;
    }
  return fail(ctx, ZKMSM_ERR_INVALID_ARG, "unknown option '%s'", name);
}

extern "C" int zkmsm_host_alloc(size_t bytes, void** out) {
  if (!out) return ZKMSM_ERR_INVALID_ARG;
  return cudaMallocHost(out, bytes ? bytes : 16) == cudaSuccess ? ZKMSM_OK : ZKMSM_ERR_CUDA;
}
extern "C" int zkmsm_host_free(void* p) { return cudaFreeHost(p) == cudaSuccess ? ZKMSM_OK : ZKMSM_ERR_CUDA; }

extern "C" int zkmsm_last_launch_count(const zkmsm_ctx* ctx) { return ctx ? ctx->last_launches : 0; }

extern "C" int zkmsm_profile(zkmsm_ctx* ctx, int enable) {
  if (!ctx) return ZKMSM_ERR_INVALID_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  if (enable && !ctx->prof) {
    ctx->prof = new (std::nothrow) LaunchProfile();
    if (!ctx->prof) return fail(ctx, ZKMSM_ERR_NOMEM, "host allocation");
    for (int i = 0; i < LaunchProfile::MAX; i++) { CU(ctx, cudaEventCreate(&ctx->prof->beg[i])); CU(ctx, cudaEventCreate(&ctx->prof->end[i])); }
  } else if (!enable && ctx->prof) {
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    for (int i = 0; i < LaunchProfile::MAX; i++) { cudaEventDestroy(ctx->prof->beg[i]); cudaEventDestroy(ctx->prof->end[i]); }
    delete ctx->prof;
    ctx->prof = nullptr;
  }
  return ZKMSM_OK;
}

extern "C" int zkmsm_profile_read(zkmsm_ctx* ctx, int max_entries, char* names, size_t name_stride, float* ms,
                                  uint32_t* threads) {
  if (!ctx || !ctx->prof) return fail(ctx, ZKMSM_ERR_INVALID_ARG, "profiling is not enabled");
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  int n = ctx->prof->n < max_entries ? ctx->prof->n : max_entries;
  for (int i = 0; i < n; i++) {
    float t = 0;
    CU(ctx, cudaEventElapsedTime(&t, ctx->prof->beg[i], ctx->prof->end[i]));
    if (ms) ms[i] = t;
    if (threads) threads[i] = ctx->prof->threads[i];
    if (names && name_stride) { strncpy(names + i * name_stride, ctx->prof->names[i], name_stride - 1); names[i * name_stride + name_stride - 1] = 0; }
  }
  return n;
}

// ------------------------------------------------------------------------------------------------
// point sets
template <class C>
static int finish_point_set(zkmsm_ctx* ctx, zkmsm_points* ps, CudaExec& ex, unsigned flags = 0) {
  typedef typename C::F F;
  if ((flags & ZKMSM_CHECK_SUBGROUP) && ps->n > 0) {   // on slab 0, before the other slabs are derived from it
    CU(ctx, cudaMemsetAsync(&ctx->d_res->aux_err, 0, sizeof(uint32_t), ctx->stream));
    ex.template launch<SubgroupCheck<C>>((uint32_t)ps->n, (uint32_t)ps->n, (const Affine<F>*)ps->d_pts, &ctx->d_res->aux_err);
    if (ex.err != cudaSuccess) return fail(ctx, ZKMSM_ERR_CUDA, "subgroup check: %s", cudaGetErrorString(ex.err));
    uint32_t bad = 0;
    CU(ctx, cudaMemcpyAsync(&bad, &ctx->d_res->aux_err, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    if (bad) return fail(ctx, ZKMSM_ERR_NOT_IN_SUBGROUP, "%u of %zu points are not in the order-r subgroup", bad, ps->n);
  }
  if (ps->precomp && ps->n > 0)
    ex.template launch<PrecomputeSlabs<C>>((uint32_t)ps->n, (uint32_t)ps->n, (uint32_t)ps->n, ps->c, ps->W,
                                           msm_wide_windows(ps->c, ps->W, ps->half, true), (Affine<F>*)ps->d_pts);
  if (ex.err != cudaSuccess) return fail(ctx, ZKMSM_ERR_CUDA, "point set kernels: %s", cudaGetErrorString(ex.err));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return ZKMSM_OK;
}

template <class C>
static int alloc_point_set(zkmsm_ctx* ctx, size_t n, unsigned flags, int curve, zkmsm_points** out) {
  typedef typename C::F F;
  if (!ctx || !out || n > (1u << 26)) return fail(ctx, ZKMSM_ERR_INVALID_ARG, "bad argument");
  *out = nullptr;
  CU(ctx, cudaSetDevice(ctx->device));
  zkmsm_points* ps = new (std::nothrow) zkmsm_points();
  if (!ps) return fail(ctx, ZKMSM_ERR_NOMEM, "host allocation");
  ps->uid = g_next_ps_uid++;
  ps->curve = curve;
  ps->device = ctx->device;
  ps->n = n;
  ps->precomp = (flags & ZKMSM_PRECOMPUTE) != 0 && n > 0;
  ps->half = (flags & ZKMSM_SUBGROUP) != 0;
  ps->c = ps->precomp ? (ctx->window_override ? ctx->window_override : msm_pick_c((uint32_t)n, true, ps->half)) : 0;
  ps->W = ps->precomp ? msm_windows(ps->c, ps->half) : 1;
  size_t slabs = ps->precomp ? ps->W : 1;
  if ((uint64_t)slabs * n >= (1ull << 31)) { delete ps; return fail(ctx, ZKMSM_ERR_INVALID_ARG, "point set too large"); }
  size_t bytes = sizeof(Affine<F>) * slabs * (n ? n : 1);
  cudaError_t e = cudaMalloc(&ps->d_pts, bytes);
  if (e != cudaSuccess) { delete ps; return fail(ctx, ZKMSM_ERR_NOMEM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e)); }
  *out = ps;
  return ZKMSM_OK;
}

template <class C>
static int load_points_impl(zkmsm_ctx* ctx, const uint32_t* xy, const uint8_t* inf, size_t n, unsigned flags, int curve,
                            zkmsm_points** out) {
  typedef typename C::F F;
  if (n > 0 && !xy) return fail(ctx, ZKMSM_ERR_INVALID_ARG, "null points");
  int rc = alloc_point_set<C>(ctx, n, flags, curve, out);
  if (rc) return rc;
  zkmsm_points* ps = *out;
  if (n == 0) return ZKMSM_OK;
  size_t cbytes = sizeof(uint32_t) * C::AFF_LIMBS * n;
  rc = ws_reserve(ctx, WS_MISC, cbytes + n);
  if (rc) { zkmsm_points_free(ctx, ps); *out = nullptr; return rc; }
  uint32_t* d_canon = (uint32_t*)ctx->ws[WS_MISC];
  uint8_t* d_inf = inf ? (uint8_t*)ctx->ws[WS_MISC] + cbytes : nullptr;
  CU(ctx, cudaMemcpyAsync(d_canon, xy, cbytes, cudaMemcpyHostToDevice, ctx->stream));
  if (inf) CU(ctx, cudaMemcpyAsync(d_inf, inf, n, cudaMemcpyHostToDevice, ctx->stream));
  CudaExec ex(ctx->stream);
  ex.template launch<LoadPoints<C>>((uint32_t)n, (uint32_t)n, (const uint32_t*)d_canon, (const uint8_t*)d_inf,
                                    (Affine<F>*)ps->d_pts);
  rc = finish_point_set<C>(ctx, ps, ex, flags);
  if (rc) { zkmsm_points_free(ctx, ps); *out = nullptr; }
  return rc;
}

extern "C" int zkmsm_g1_load_points(zkmsm_ctx* ctx, const uint32_t* xy, const uint8_t* inf, size_t n, unsigned flags,
                                    zkmsm_points** out) {
  return load_points_impl<G1>(ctx, xy, inf, n, flags, 1, out);
}
extern "C" int zkmsm_g2_load_points(zkmsm_ctx* ctx, const uint32_t* xy, const uint8_t* inf, size_t n, unsigned flags,
                                    zkmsm_points** out) {
  return load_points_impl<G2>(ctx, xy, inf, n, flags, 2, out);
}

extern "C" int zkmsm_points_free(zkmsm_ctx* ctx, zkmsm_points* ps) {
  if (!ps) return ZKMSM_ERR_INVALID_ARG;
  cudaSetDevice(ps->device);
  if (ctx) { cudaStreamSynchronize(ctx->stream); graphs_drop(ctx, ps->uid); }
  if (ps->d_pts) cudaFree(ps->d_pts);
  delete ps;
  return ZKMSM_OK;
}

extern "C" size_t zkmsm_points_len(const zkmsm_points* ps) { return ps ? ps->n : 0; }

extern "C" int zkmsm_points_info(const zkmsm_points* ps, unsigned* c, unsigned* windows, int* precomputed, int* subgroup) {
  if (!ps) return ZKMSM_ERR_INVALID_ARG;
  if (c) *c = ps->c;
  if (windows) *windows = ps->precomp ? ps->W : 0;
  if (precomputed) *precomputed = ps->precomp ? 1 : 0;
  if (subgroup) *subgroup = ps->half ? 1 : 0;
  return ZKMSM_OK;
}

template <class C>
static int points_read_impl(zkmsm_ctx* ctx, const zkmsm_points* ps, size_t first, size_t n, uint32_t* xy, uint8_t* inf) {
  typedef typename C::F F;
  if (!ctx || !ps || first + n > ps->n || (n && !xy)) return fail(ctx, ZKMSM_ERR_INVALID_ARG, "bad argument");
  if (n == 0) return ZKMSM_OK;
  CU(ctx, cudaSetDevice(ctx->device));
  size_t cbytes = sizeof(uint32_t) * C::AFF_LIMBS * n;
  int rc = ws_reserve(ctx, WS_MISC, cbytes + n);
  if (rc) return rc;
  uint32_t* d_canon = (uint32_t*)ctx->ws[WS_MISC];
  uint8_t* d_inf = (uint8_t*)ctx->ws[WS_MISC] + cbytes;
  CudaExec ex(ctx->stream);
  ex.template launch<StorePoints<C>>((uint32_t)n, (uint32_t)n, (const Affine<F>*)ps->d_pts + first, d_canon, d_inf);
  if (ex.err != cudaSuccess) return fail(ctx, ZKMSM_ERR_CUDA, "StorePoints: %s", cudaGetErrorString(ex.err));
  CU(ctx, cudaMemcpyAsync(xy, d_canon, cbytes, cudaMemcpyDeviceToHost, ctx->stream));
  if (inf) CU(ctx, cudaMemcpyAsync(inf, d_inf, n, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return ZKMSM_OK;
}
extern "C" int zkmsm_points_read(zkmsm_ctx* ctx, const zkmsm_points* ps, size_t first, size_t n, uint32_t* xy,
                                 uint8_t* inf) {
  if (!ps) return ZKMSM_ERR_INVALID_ARG;
  return ps->curve == 1 ? points_read_impl<G1>(ctx, ps, first, n, xy, inf) : points_read_impl<G2>(ctx, ps, first, n, xy, inf);
}

// ------------------------------------------------------------------------------------------------
// MSM
// rank/world: bucket-range split of a precomputed set (world = 1: the whole MSM)
template <class C>
static int msm_enqueue_impl(zkmsm_ctx* ctx, const zkmsm_points* ps, const uint32_t* d_scalars, size_t n, int curve,
                            bool want_affine, uint32_t* d_partial_out, uint32_t rank = 0, uint32_t world = 1) {
  typedef typename C::F F;
  if (!ctx || !ps) return fail(ctx, ZKMSM_ERR_INVALID_ARG, "null argument");
  if (ps->curve != curve) return fail(ctx, ZKMSM_ERR_INVALID_ARG, "point set is for the other group");
  if (ps->device != ctx->device) return fail(ctx, ZKMSM_ERR_INVALID_ARG, "point set lives on another device");
  if (n > ps->n) return fail(ctx, ZKMSM_ERR_TOO_FEW_POINTS, "%zu scalars but only %zu points", n, ps->n);
  if (n > 0 && !d_scalars) return fail(ctx, ZKMSM_ERR_INVALID_ARG, "null scalars");
  if (world == 0 || rank >= world || (world & (world - 1)) != 0)
    return fail(ctx, ZKMSM_ERR_INVALID_ARG, "bucket-range split: world must be a power of two and rank < world");
  if (world > 1 && !ps->precomp) return fail(ctx, ZKMSM_ERR_INVALID_ARG, "bucket-range split needs a ZKMSM_PRECOMPUTE point set");
  CU(ctx, cudaSetDevice(ctx->device));
  ctx->pending_words = C::AFF_LIMBS;
  ctx->last_launches = 0;
  XYZZ<F>* d_xyzz = d_partial_out ? (XYZZ<F>*)d_partial_out : (XYZZ<F>*)ctx->d_res->xyzz;
  if (n == 0) {  // empty sum = AtInfinity (polynomial.rs:276)
    ctx->pending = 2;
    if (d_partial_out) CU(ctx, cudaMemsetAsync(d_partial_out, 0, sizeof(XYZZ<F>), ctx->stream));
    return ZKMSM_OK;
  }
  MsmTuning tune = ctx->tune;
  // G2 keeps the chunked accumulation (+ the one-launch fix-up): its per-bucket kernel is bound by registers and by
  // the spread of the bucket sizes without pre-reduction rounds (measured 7.9 against 4.7 ms at 2^18)
  if (curve != 1 && tune.acc_G == 0) tune.no_bucket_acc = 1;
  // the G2 batched-affine kernel keeps two blocks of 128 threads per SM (255 registers), not three
  if (curve != 1 && tune.batch_blocks == 0) tune.batch_blocks = 2;
  unsigned c = ps->precomp ? ps->c : (ctx->window_override ? ctx->window_override : msm_pick_c((uint32_t)n, false, ps->half));
  if (!msm_fits(n, c, ps->half))
    return fail(ctx, ZKMSM_ERR_INVALID_ARG, "%zu terms at window %u exceed 2^32 sorted pairs; use a wider window", n, c);
  if (world > 1 && world > (1u << (c - 1))) return fail(ctx, ZKMSM_ERR_INVALID_ARG, "bucket-range split: more ranks than buckets");
  MsmPlan p = msm_plan((uint32_t)n, c, ps->precomp, (uint32_t)ps->n, ps->half, true, (uint32_t)ctx->sms * 256u, tune, rank, world);
  // G2 takes the rounds too since its kernel has the field arithmetic inlined (tu_g2_batch.cu: 1.1 ns per affine
  // addition against 1.2 ns per mixed addition, and far fewer chunk sums to fix up: 6.64 -> 6.14 ms at 2^18)
  if (curve != 1 && !tune.batch_g2) p.batch_rounds = 0;
  else if (tune.batch_rounds < 0) {
    p.batch_rounds = msm_default_batch_rounds(p, tune);
    // (G2 gains per addition are small; below three rounds' worth of work the rounds' fixed costs win: a half-size
    // share at 2^18 measured 3.93 ms with two rounds against 3.74 ms without)
    if (curve != 1 && p.batch_rounds < 3) p.batch_rounds = 0;
  }
  int rc;
  uint32_t nseg = (p.nb + SCAN_SEG - 1) / SCAN_SEG + 1;
  if ((rc = ws_reserve(ctx, WS_HIST, sizeof(uint32_t) * p.nb)) || (rc = ws_reserve(ctx, WS_OFFSETS, sizeof(uint32_t) * (p.nb + 1))) ||
      (rc = ws_reserve(ctx, WS_SEGSUM, sizeof(uint32_t) * nseg)) || (rc = ws_reserve(ctx, WS_ENTRIES, sizeof(Entry) * (size_t)p.max_entries)) ||
      (rc = ws_reserve(ctx, WS_BUCKETS, sizeof(XYZZ<F>) * (size_t)p.nb)) ||
      (rc = ws_reserve(ctx, WS_PARTIALS, sizeof(XYZZ<F>) * msm_partial_slots(p))) ||
      (rc = ws_reserve(ctx, WS_PKEYS, sizeof(uint32_t) * msm_partial_slots(p))) ||
      (rc = ws_reserve(ctx, WS_REDUCED, sizeof(XYZZ<F>) * msm_reduced_slots(p))))
    return rc;
  MsmBuffers<C> b;
  memset(&b, 0, sizeof(b));
  if (p.batch_rounds > 0) {
    size_t pre_n = msm_pre_slots(p);
    if ((rc = ws_reserve(ctx, WS_PRE_A, sizeof(Affine<F>) * pre_n)) || (rc = ws_reserve(ctx, WS_PRE_B, sizeof(Affine<F>) * pre_n)) ||
        (rc = ws_reserve(ctx, WS_PRE_PREFIX, sizeof(F) * 2 * msm_prefix_slots(p))) || (rc = ws_reserve(ctx, WS_PRE_OFF, sizeof(uint32_t) * 3 * ((size_t)p.nb + 4))) ||
        (rc = ws_reserve(ctx, WS_PRE_ENTRIES, sizeof(Entry) * pre_n))) {
      // the rounds are an optimisation: without room for their scratch the MSM runs on the XYZZ accumulation alone
      // (forced rounds still report the failure)
      if (rc != ZKMSM_ERR_NOMEM || tune.batch_rounds >= 0) return rc;
      for (int slot : {WS_PRE_A, WS_PRE_B, WS_PRE_PREFIX, WS_PRE_ENTRIES})
        if (ctx->ws[slot]) { cudaFree(ctx->ws[slot]); ctx->ws[slot] = nullptr; ctx->ws_bytes[slot] = 0; }
      strcpy(ctx->err, "ok");
      p.batch_rounds = 0;
    }
  }
  if (p.batch_rounds > 0) {
    b.pre_pts[0] = (Affine<F>*)ctx->ws[WS_PRE_A];
    b.pre_pts[1] = (Affine<F>*)ctx->ws[WS_PRE_B];
    b.pre_prefix = (F*)ctx->ws[WS_PRE_PREFIX];
    b.pre_off[0] = (uint32_t*)ctx->ws[WS_PRE_OFF];
    b.pre_off[1] = b.pre_off[0] + p.nb + 4;
    b.pre_cnt = b.pre_off[1] + p.nb + 4;
    b.pre_entries = (Entry*)ctx->ws[WS_PRE_ENTRIES];
  }
  b.hist_cursor = (uint32_t*)ctx->ws[WS_HIST];
  b.offsets = (uint32_t*)ctx->ws[WS_OFFSETS];
  b.segsum = (uint32_t*)ctx->ws[WS_SEGSUM];
  b.entries = (Entry*)ctx->ws[WS_ENTRIES];
  b.bucket_sums = (XYZZ<F>*)ctx->ws[WS_BUCKETS];
  b.partials = (XYZZ<F>*)ctx->ws[WS_PARTIALS];
  b.partial_keys = (uint32_t*)ctx->ws[WS_PKEYS];
  b.reduced = (XYZZ<F>*)ctx->ws[WS_REDUCED];
  b.err = &ctx->d_res->err;
  b.big = &ctx->d_res->big;
  auto enqueue = [&](CudaExec& ex) {
    msm_launch<C>(ex, p, tune, b, (const Affine<F>*)ps->d_pts, d_scalars, want_affine ? (XYZZ<F>*)nullptr : d_xyzz,
                  want_affine ? ctx->d_res->affine : (uint32_t*)nullptr, want_affine ? &ctx->d_res->inf : (uint32_t*)nullptr);
    // a partial that leaves the context without a status read carries its error flag inside the blob
    if (d_partial_out) ex.template launch<PoisonPartial<C>>(1u, (const uint32_t*)b.err, d_xyzz);
  };
  // Replay a captured CUDA graph of the sequence when nothing it depends on changed (the ~30-40 launches then cost
  // one submission and run back to back); per-launch profiling and ZKMSM_NO_GRAPH launch kernel by kernel.
  if (!ctx->prof && !tune.no_graph) {
    GraphKey key{ps->uid, ctx->ws_epoch, ctx->tune_epoch, n, d_scalars, d_partial_out, curve | (want_affine ? 16 : 0), rank, world};
    GraphSlot* slot = nullptr;
    GraphSlot* victim = &ctx->graphs[0];
    for (int i = 0; i < kGraphSlots; i++) {
      GraphSlot& g = ctx->graphs[i];
      if (g.exec && g.key == key) { slot = &g; break; }
      if (!g.exec || (victim->exec && g.used < victim->used)) victim = &g;
    }
    if (!slot && cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      CudaExec ex(ctx->stream, nullptr, tune.no_coop != 0, tune.ntt_no_fuse != 0);
      enqueue(ex);
      cudaGraph_t graph = nullptr;
      cudaError_t e = cudaStreamEndCapture(ctx->stream, &graph);
      if (ex.err != cudaSuccess || e != cudaSuccess) {
        if (graph) cudaGraphDestroy(graph);
        cudaGetLastError();
        return fail(ctx, ZKMSM_ERR_CUDA, "msm capture: %s", cudaGetErrorString(ex.err != cudaSuccess ? ex.err : e));
      }
      if (victim->exec) cudaGraphExecDestroy(victim->exec);
      victim->exec = nullptr;
      e = cudaGraphInstantiate(&victim->exec, graph, 0);
      cudaGraphDestroy(graph);
      if (e != cudaSuccess) { victim->exec = nullptr; return fail(ctx, ZKMSM_ERR_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e)); }
      victim->key = key;
      victim->launches = ex.launches;
      slot = victim;
    } else if (!slot) {
      cudaGetLastError();   // the stream cannot be captured (e.g. the legacy default stream): plain launches below
    }
    if (slot) {
      slot->used = ++ctx->graph_clock;
      CU(ctx, cudaGraphLaunch(slot->exec, ctx->stream));
      ctx->last_launches = slot->launches;
      ctx->pending = 1;
      return ZKMSM_OK;
    }
  }
  if (ctx->prof) ctx->prof->n = 0;
  CudaExec ex(ctx->stream, ctx->prof, tune.no_coop != 0, tune.ntt_no_fuse != 0);
  enqueue(ex);
  ctx->last_launches = ex.launches;
  if (ex.err != cudaSuccess) return fail(ctx, ZKMSM_ERR_CUDA, "msm launch: %s", cudaGetErrorString(ex.err));
  ctx->pending = 1;
  return ZKMSM_OK;
}

// wait for the enqueued MSM; copies the result block to pinned host memory
static int msm_collect(zkmsm_ctx* ctx) {
  if (ctx->pending == 0) return fail(ctx, ZKMSM_ERR_INVALID_ARG, "no MSM pending");
  if (ctx->pending == 2) {
    memset(ctx->h_res, 0, sizeof(ResultBlock));
    ctx->h_res->inf = 1;
    CU(ctx, cudaStreamSynchronize(ctx->stream));
    ctx->pending = 0;
    return ZKMSM_OK;
  }
  CU(ctx, cudaMemcpyAsync(ctx->h_res, ctx->d_res, sizeof(ResultBlock), cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->pending = 0;
  if (ctx->h_res->err & ERR_SCALAR_RANGE) return fail(ctx, ZKMSM_ERR_SCALAR_RANGE, "a scalar is out of range (bit 255 set, or >= r for a ZKMSM_SUBGROUP point set)");
  return ZKMSM_OK;
}

static int msm_result_impl(zkmsm_ctx* ctx, int words, uint32_t* out_xy, int* out_is_inf) {
  if (!ctx || !out_xy || !out_is_inf) return fail(ctx, ZKMSM_ERR_INVALID_ARG, "null argument");
  if (ctx->pending && ctx->pending_words != words) return fail(ctx, ZKMSM_ERR_INVALID_ARG, "pending MSM is for the other group");
  CU(ctx, cudaSetDevice(ctx->device));
  int rc = msm_collect(ctx);
  if (rc) return rc;
  *out_is_inf = ctx->h_res->inf ? 1 : 0;
  if (ctx->h_res->inf) memset(out_xy, 0, sizeof(uint32_t) * words);
  else memcpy(out_xy, ctx->h_res->affine, sizeof(uint32_t) * words);
  return ZKMSM_OK;
}

static int upload_scalars(zkmsm_ctx* ctx, const uint32_t* scalars, size_t n, const uint32_t** d_out) {
  *d_out = nullptr;
  if (n == 0) return ZKMSM_OK;
  if (!scalars) return fail(ctx, ZKMSM_ERR_INVALID_ARG, "null scalars");
  CU(ctx, cudaSetDevice(ctx->device));
  int rc = ws_reserve(ctx, WS_SCALARS, sizeof(uint32_t) * SCALAR_LIMBS * n);
  if (rc) return rc;
  CU(ctx, cudaMemcpyAsync(ctx->ws[WS_SCALARS], scalars, sizeof(uint32_t) * SCALAR_LIMBS * n, cudaMemcpyHostToDevice, ctx->stream));
  *d_out = (const uint32_t*)ctx->ws[WS_SCALARS];
  return ZKMSM_OK;
}

template <class C>
static int msm_host_impl(zkmsm_ctx* ctx, const zkmsm_points* ps, const uint32_t* scalars, size_t n, int curve,
                         uint32_t* out_xy, int* out_is_inf) {
  if (!ctx || !ps || !out_xy || !out_is_inf) return fail(ctx, ZKMSM_ERR_INVALID_ARG, "null argument");
  if (n > ps->n) return fail(ctx, ZKMSM_ERR_TOO_FEW_POINTS, "%zu scalars but only %zu points", n, ps->n);
  const uint32_t* d_s;
  int rc = upload_scalars(ctx, scalars, n, &d_s);
  if (rc) return rc;
  rc = msm_enqueue_impl<C>(ctx, ps, d_s, n, curve, true, nullptr);
  if (rc) return rc;
  return msm_result_impl(ctx, C::AFF_LIMBS, out_xy, out_is_inf);
}

// asynchronous first half of msm_host_impl: copy the scalars (truly asynchronous from pinned memory) and enqueue
template <class C>
static int msm_begin_impl(zkmsm_ctx* ctx, const zkmsm_points* ps, const uint32_t* scalars, size_t n, int curve) {
  if (!ctx || !ps) return fail(ctx, ZKMSM_ERR_INVALID_ARG, "null argument");
  if (n > ps->n) return fail(ctx, ZKMSM_ERR_TOO_FEW_POINTS, "%zu scalars but only %zu points", n, ps->n);
  const uint32_t* d_s;
  int rc = upload_scalars(ctx, scalars, n, &d_s);
  if (rc) return rc;
  return msm_enqueue_impl<C>(ctx, ps, d_s, n, curve, true, nullptr);
}
extern "C" int zkmsm_g1_msm_begin(zkmsm_ctx* ctx, const zkmsm_points* ps, const uint32_t* s, size_t n) {
  return msm_begin_impl<G1>(ctx, ps, s, n, 1);
}
extern "C" int zkmsm_g2_msm_begin(zkmsm_ctx* ctx, const zkmsm_points* ps, const uint32_t* s, size_t n) {
  return msm_begin_impl<G2>(ctx, ps, s, n, 2);
}

extern "C" int zkmsm_g1_msm(zkmsm_ctx* ctx, const zkmsm_points* ps, const uint32_t* s, size_t n, uint32_t* out, int* inf) {
  return msm_host_impl<G1>(ctx, ps, s, n, 1, out, inf);
}
extern "C" int zkmsm_g2_msm(zkmsm_ctx* ctx, const zkmsm_points* ps, const uint32_t* s, size_t n, uint32_t* out, int* inf) {
  return msm_host_impl<G2>(ctx, ps, s, n, 2, out, inf);
}
extern "C" int zkmsm_g1_msm_enqueue(zkmsm_ctx* ctx, const zkmsm_points* ps, const uint32_t* ds, size_t n) {
  return msm_enqueue_impl<G1>(ctx, ps, ds, n, 1, true, nullptr);
}
extern "C" int zkmsm_g2_msm_enqueue(zkmsm_ctx* ctx, const zkmsm_points* ps, const uint32_t* ds, size_t n) {
  return msm_enqueue_impl<G2>(ctx, ps, ds, n, 2, true, nullptr);
}
extern "C" int zkmsm_g1_msm_result(zkmsm_ctx* ctx, uint32_t* out, int* inf) { return msm_result_impl(ctx, 24, out, inf); }
extern "C" int zkmsm_g2_msm_result(zkmsm_ctx* ctx, uint32_t* out, int* inf) { return msm_result_impl(ctx, 48, out, inf); }
extern "C" int zkmsm_g1_msm_device(zkmsm_ctx* ctx, const zkmsm_points* ps, const uint32_t* ds, size_t n, uint32_t* out, int* inf) {
  int rc = msm_enqueue_impl<G1>(ctx, ps, ds, n, 1, true, nullptr);
  return rc ? rc : msm_result_impl(ctx, 24, out, inf);
}
extern "C" int zkmsm_g2_msm_device(zkmsm_ctx* ctx, const zkmsm_points* ps, const uint32_t* ds, size_t n, uint32_t* out, int* inf) {
  int rc = msm_enqueue_impl<G2>(ctx, ps, ds, n, 2, true, nullptr);
  return rc ? rc : msm_result_impl(ctx, 48, out, inf);
}

template <class C>
static int oneshot_impl(zkmsm_ctx* ctx, const uint32_t* xy, const uint8_t* inf, const uint32_t* scalars, size_t n, int curve,
                        uint32_t* out_xy, int* out_is_inf) {
  zkmsm_points* ps = nullptr;
  int rc = load_points_impl<C>(ctx, xy, inf, n, 0, curve, &ps);
  if (rc) return rc;
  rc = msm_host_impl<C>(ctx, ps, scalars, n, curve, out_xy, out_is_inf);
  zkmsm_points_free(ctx, ps);
  return rc;
}
extern "C" int zkmsm_g1_msm_oneshot(zkmsm_ctx* ctx, const uint32_t* xy, const uint8_t* inf, const uint32_t* s, size_t n,
                                    uint32_t* out, int* oinf) {
  return oneshot_impl<G1>(ctx, xy, inf, s, n, 1, out, oinf);
}
extern "C" int zkmsm_g2_msm_oneshot(zkmsm_ctx* ctx, const uint32_t* xy, const uint8_t* inf, const uint32_t* s, size_t n,
                                    uint32_t* out, int* oinf) {
  return oneshot_impl<G2>(ctx, xy, inf, s, n, 2, out, oinf);
}

// ---- partials / combine
template <class C>
static int partial_impl(zkmsm_ctx* ctx, const zkmsm_points* ps, const uint32_t* scalars, size_t n, int curve, uint32_t* out,
                        uint32_t rank = 0, uint32_t world = 1) {
  typedef typename C::F F;
  if (!ctx || !ps || !out) return fail(ctx, ZKMSM_ERR_INVALID_ARG, "null argument");
  if (n > ps->n) return fail(ctx, ZKMSM_ERR_TOO_FEW_POINTS, "%zu scalars but only %zu points", n, ps->n);
  const uint32_t* d_s;
  int rc = upload_scalars(ctx, scalars, n, &d_s);
  if (rc) return rc;
  rc = msm_enqueue_impl<C>(ctx, ps, d_s, n, curve, false, nullptr, rank, world);
  if (rc) return rc;
  rc = msm_collect(ctx);
  if (rc) return rc;
  memcpy(out, ctx->h_res->xyzz, sizeof(XYZZ<F>));  // n == 0: zeros = infinity
  return ZKMSM_OK;
}
// stream-ordered: the partial stays in the caller's device buffer; an out-of-range scalar poisons the blob and the
// combine step reports ZKMSM_ERR_SCALAR_RANGE (PoisonPartial, msm.cuh)
template <class C>
static int partial_device_impl(zkmsm_ctx* ctx, const zkmsm_points* ps, const uint32_t* ds, size_t n, int curve, uint32_t* d_out,
                               uint32_t rank = 0, uint32_t world = 1) {
  if (!d_out) return fail(ctx, ZKMSM_ERR_INVALID_ARG, "null argument");
  int rc = msm_enqueue_impl<C>(ctx, ps, ds, n, curve, false, d_out, rank, world);
  if (rc) return rc;
  ctx->pending = 0;  // result lives in the caller's buffer, stream-ordered
  return ZKMSM_OK;
}
extern "C" int zkmsm_g1_msm_partial(zkmsm_ctx* ctx, const zkmsm_points* ps, const uint32_t* s, size_t n, uint32_t* out) {
  return partial_impl<G1>(ctx, ps, s, n, 1, out);
}
extern "C" int zkmsm_g2_msm_partial(zkmsm_ctx* ctx, const zkmsm_points* ps, const uint32_t* s, size_t n, uint32_t* out) {
  return partial_impl<G2>(ctx, ps, s, n, 2, out);
}
extern "C" int zkmsm_g1_msm_partial_device(zkmsm_ctx* ctx, const zkmsm_points* ps, const uint32_t* ds, size_t n, uint32_t* d_out) {
  return partial_device_impl<G1>(ctx, ps, ds, n, 1, d_out);
}
extern "C" int zkmsm_g2_msm_partial_device(zkmsm_ctx* ctx, const zkmsm_points* ps, const uint32_t* ds, size_t n, uint32_t* d_out) {
  return partial_device_impl<G2>(ctx, ps, ds, n, 2, d_out);
}
extern "C" int zkmsm_g1_msm_partial_range(zkmsm_ctx* ctx, const zkmsm_points* ps, const uint32_t* s, size_t n, unsigned rank,
                                          unsigned world, uint32_t* out) {
  return partial_impl<G1>(ctx, ps, s, n, 1, out, rank, world);
}
extern "C" int zkmsm_g2_msm_partial_range(zkmsm_ctx* ctx, const zkmsm_points* ps, const uint32_t* s, size_t n, unsigned rank,
                                          unsigned world, uint32_t* out) {
  return partial_impl<G2>(ctx, ps, s, n, 2, out, rank, world);
}
extern "C" int zkmsm_g1_msm_partial_range_device(zkmsm_ctx* ctx, const zkmsm_points* ps, const uint32_t* ds, size_t n,
                                                 unsigned rank, unsigned world, uint32_t* d_out) {
  return partial_device_impl<G1>(ctx, ps, ds, n, 1, d_out, rank, world);
}
extern "C" int zkmsm_g2_msm_partial_range_device(zkmsm_ctx* ctx, const zkmsm_points* ps, const uint32_t* ds, size_t n,
                                                 unsigned rank, unsigned world, uint32_t* d_out) {
  return partial_device_impl<G2>(ctx, ps, ds, n, 2, d_out, rank, world);
}

// one combine kernel: k partials, stride_words apart, summed in order and converted to canonical affine; a poisoned
// partial raises d_res->err
template <class C>
static void combine_launch(zkmsm_ctx* ctx, CudaExec& ex, const XYZZ<typename C::F>* d_parts, uint32_t k, uint32_t stride_words,
                           uint32_t* out_affine, uint32_t* out_inf) {
  if (std::is_same<C, G1>::value && k > 2 && k <= 32 && !ctx->tune.no_coop)
    ex.timed("combine_partials", k, 1, [&] { return zk_coop_combine_g1(ctx->stream, k, (const XYZZ<Fp>*)d_parts, stride_words, out_affine, out_inf, &ctx->d_res->err); });
  else if (std::is_same<C, G2>::value && k > 1 && k <= 32 && !ctx->tune.no_coop)   // a lone thread needs ~0.1 ms per G2 addition
    ex.timed("combine_partials", k, 1, [&] { return zk_coop_combine_g2(ctx->stream, k, (const XYZZ<Fp2>*)d_parts, stride_words, out_affine, out_inf, &ctx->d_res->err); });
  else
    ex.template launch<CombinePartials<C>>(1u, k, d_parts, stride_words, out_affine, out_inf, &ctx->d_res->err);
}

// enqueue_only: stream-ordered, the result is fetched later with zkmsm_g{1,2}_msm_result
template <class C>
static int combine_impl(zkmsm_ctx* ctx, const uint32_t* parts, bool on_device, size_t k, uint32_t* out_xy, int* out_is_inf,
                        bool enqueue_only = false) {
  typedef typename C::F F;
  if (!ctx || (k && !parts) || (!enqueue_only && (!out_xy || !out_is_inf)) || k > 4096) return fail(ctx, ZKMSM_ERR_INVALID_ARG, "bad argument");
  CU(ctx, cudaSetDevice(ctx->device));
  const XYZZ<F>* d_parts = (const XYZZ<F>*)parts;
  if (!on_device && k) {
    int rc = ws_reserve(ctx, WS_MISC, sizeof(XYZZ<F>) * k);
    if (rc) return rc;
    CU(ctx, cudaMemcpyAsync(ctx->ws[WS_MISC], parts, sizeof(XYZZ<F>) * k, cudaMemcpyHostToDevice, ctx->stream));
    d_parts = (const XYZZ<F>*)ctx->ws[WS_MISC];
  }
  CU(ctx, cudaMemsetAsync(&ctx->d_res->err, 0, sizeof(uint32_t), ctx->stream));
  CudaExec ex(ctx->stream, nullptr, ctx->tune.no_coop != 0);
  combine_launch<C>(ctx, ex, d_parts, (uint32_t)k, (uint32_t)(sizeof(XYZZ<F>) / 4), ctx->d_res->affine, &ctx->d_res->inf);
  if (ex.err != cudaSuccess) return fail(ctx, ZKMSM_ERR_CUDA, "combine: %s", cudaGetErrorString(ex.err));
  ctx->pending = 1;
  ctx->pending_words = C::AFF_LIMBS;
  if (enqueue_only) return ZKMSM_OK;
  return msm_result_impl(ctx, C::AFF_LIMBS, out_xy, out_is_inf);
}
extern "C" int zkmsm_g1_combine(zkmsm_ctx* ctx, const uint32_t* parts, size_t k, uint32_t* out, int* inf) {
  return combine_impl<G1>(ctx, parts, false, k, out, inf);
}
extern "C" int zkmsm_g2_combine(zkmsm_ctx* ctx, const uint32_t* parts, size_t k, uint32_t* out, int* inf) {
  return combine_impl<G2>(ctx, parts, false, k, out, inf);
}
extern "C" int zkmsm_g1_combine_device(zkmsm_ctx* ctx, const uint32_t* parts, size_t k, uint32_t* out, int* inf) {
  return combine_impl<G1>(ctx, parts, true, k, out, inf);
}
extern "C" int zkmsm_g2_combine_device(zkmsm_ctx* ctx, const uint32_t* parts, size_t k, uint32_t* out, int* inf) {
  return combine_impl<G2>(ctx, parts, true, k, out, inf);
}
extern "C" int zkmsm_g1_combine_enqueue(zkmsm_ctx* ctx, const uint32_t* parts_device, size_t k) {
  return combine_impl<G1>(ctx, parts_device, true, k, nullptr, nullptr, true);
}
extern "C" int zkmsm_g2_combine_enqueue(zkmsm_ctx* ctx, const uint32_t* parts_device, size_t k) {
  return combine_impl<G2>(ctx, parts_device, true, k, nullptr, nullptr, true);
}

// ------------------------------------------------------------------------------------------------
// fixed-base vector scalar multiplication
template <class C>
static int base_table(zkmsm_ctx* ctx, const uint32_t* base_xy, CudaExec& ex, Affine<typename C::F>** table_out) {
  typedef typename C::F F;
  size_t need = sizeof(uint32_t) * C::AFF_LIMBS + 16 + sizeof(XYZZ<F>) * 256 + sizeof(Affine<F>) * 256;
  int rc = ws_reserve(ctx, WS_REDUCED, need);
  if (rc) return rc;
  char* base = (char*)ctx->ws[WS_REDUCED];
  XYZZ<F>* chain = (XYZZ<F>*)base;
  Affine<F>* table = (Affine<F>*)(base + sizeof(XYZZ<F>) * 256);
  uint32_t* d_base = (uint32_t*)(base + sizeof(XYZZ<F>) * 256 + sizeof(Affine<F>) * 256);
  CU(ctx, cudaMemcpyAsync(d_base, base_xy, sizeof(uint32_t) * C::AFF_LIMBS, cudaMemcpyHostToDevice, ctx->stream));
  ex.template launch<BaseTableChain<C>>(1u, (const uint32_t*)d_base, chain);
  ex.template launch<BaseTableAffine<C>>(256u, (const XYZZ<F>*)chain, table);
  *table_out = table;
  return ZKMSM_OK;
}

template <class C>
static int points_from_scalars_impl(zkmsm_ctx* ctx, const uint32_t* base_xy, const uint32_t* scalars, size_t n, unsigned flags,
                                    int curve, zkmsm_points** out) {
  typedef typename C::F F;
  if (!ctx || !base_xy || !out || (n && !scalars)) return fail(ctx, ZKMSM_ERR_INVALID_ARG, "null argument");
  int rc = alloc_point_set<C>(ctx, n, flags, curve, out);
  if (rc) return rc;
  zkmsm_points* ps = *out;
  if (n == 0) return ZKMSM_OK;
  const uint32_t* d_s;
  rc = upload_scalars(ctx, scalars, n, &d_s);
  CudaExec ex(ctx->stream);
  Affine<F>* table = nullptr;
  if (!rc) rc = base_table<C>(ctx, base_xy, ex, &table);
  if (!rc) {
    ex.template launch<FixedBaseMul<C>>((uint32_t)n, (uint32_t)n, d_s, (const Affine<F>*)table, (Affine<F>*)ps->d_pts);
    rc = finish_point_set<C>(ctx, ps, ex);
  }
  if (rc) { zkmsm_points_free(ctx, ps); *out = nullptr; }
  return rc;
}
extern "C" int zkmsm_g1_points_from_scalars(zkmsm_ctx* ctx, const uint32_t* base, const uint32_t* s, size_t n, unsigned flags,
                                            zkmsm_points** out) {
  return points_from_scalars_impl<G1>(ctx, base, s, n, flags, 1, out);
}
extern "C" int zkmsm_g2_points_from_scalars(zkmsm_ctx* ctx, const uint32_t* base, const uint32_t* s, size_t n, unsigned flags,
                                            zkmsm_points** out) {
  return points_from_scalars_impl<G2>(ctx, base, s, n, flags, 2, out);
}

template <class C>
static int mul_base_impl(zkmsm_ctx* ctx, const uint32_t* base_xy, const uint32_t* scalars, size_t n, int curve, uint32_t* out_xy,
                         uint8_t* out_inf) {
  zkmsm_points* ps = nullptr;
  if (n && !out_xy) return fail(ctx, ZKMSM_ERR_INVALID_ARG, "null output");
  int rc = points_from_scalars_impl<C>(ctx, base_xy, scalars, n, 0, curve, &ps);
  if (rc) return rc;
  rc = zkmsm_points_read(ctx, ps, 0, n, out_xy, out_inf);
  zkmsm_points_free(ctx, ps);
  return rc;
}
extern "C" int zkmsm_g1_mul_base(zkmsm_ctx* ctx, const uint32_t* base, const uint32_t* s, size_t n, uint32_t* out, uint8_t* inf) {
  return mul_base_impl<G1>(ctx, base, s, n, 1, out, inf);
}
extern "C" int zkmsm_g2_mul_base(zkmsm_ctx* ctx, const uint32_t* base, const uint32_t* s, size_t n, uint32_t* out, uint8_t* inf) {
  return mul_base_impl<G2>(ctx, base, s, n, 2, out, inf);
}

// ------------------------------------------------------------------------------------------------
// Fr helpers for kernel ARGUMENTS only (the loop index k of t = prod (x - k) in Montgomery form): k R mod r is
// stepped by adding R mod r with one conditional subtraction of r.  No data-path arithmetic happens on the host.
static const uint32_t FR_ONE_HOST[8] = {0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau, 0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u};
static const uint32_t FR_P_HOST[8] = {0x00000001u, 0xffffffffu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
static void fr_host_add_one_mont(Fr& a) {
  uint64_t c = 0;
  uint32_t t[8], d[8];
  for (int i = 0; i < 8; i++) { uint64_t s = (uint64_t)a.v[i] + FR_ONE_HOST[i] + c; t[i] = (uint32_t)s; c = s >> 32; }
  uint64_t b = 0;
  for (int i = 0; i < 8; i++) { uint64_t s = (uint64_t)t[i] - FR_P_HOST[i] - b; d[i] = (uint32_t)s; b = (s >> 32) & 1; }
  bool ge = c || !b;
  for (int i = 0; i < 8; i++) a.v[i] = ge ? d[i] : t[i];
}

// ------------------------------------------------------------------------------------------------
// Fr witness aggregation
extern "C" int zkmsm_fr_aggregate(zkmsm_ctx* ctx, const uint32_t* polys, size_t n_wires, size_t n, const uint32_t* wires,
                                  uint32_t* out) {
  if (!ctx || !out || (n_wires && n && (!polys || !wires)) || n > (1u << 28) || n_wires > (1u << 28))
    return fail(ctx, ZKMSM_ERR_INVALID_ARG, "bad argument");
  if (n == 0) return ZKMSM_OK;
  CU(ctx, cudaSetDevice(ctx->device));
  size_t mat = sizeof(uint32_t) * 8 * n_wires * n, wb = sizeof(uint32_t) * 8 * n_wires, ob = sizeof(uint32_t) * 8 * n;
  int rc;
  if ((rc = ws_reserve(ctx, WS_ENTRIES, mat + 256)) || (rc = ws_reserve(ctx, WS_MISC, 2 * wb + ob + 256))) return rc;
  uint32_t* d_polys = (uint32_t*)ctx->ws[WS_ENTRIES];
  char* misc = (char*)ctx->ws[WS_MISC];
  uint32_t* d_wires = (uint32_t*)misc;
  Fr* d_wires_mont = (Fr*)(misc + ((wb + 255) / 256) * 256);
  uint32_t* d_out = (uint32_t*)(misc + 2 * ((wb + 255) / 256) * 256);
  if (n_wires) {
    CU(ctx, cudaMemcpyAsync(d_polys, polys, mat, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(d_wires, wires, wb, cudaMemcpyHostToDevice, ctx->stream));
  }
  CudaExec ex(ctx->stream);
  ex.template launch<FrToMont>((uint32_t)n_wires, (uint32_t)n_wires, (const uint32_t*)d_wires, d_wires_mont);
  ex.template launch<FrAggregate>((uint32_t)n, (uint32_t)n_wires, (uint32_t)n, (const Fr*)d_wires_mont, (const uint32_t*)d_polys, d_out);
  if (ex.err != cudaSuccess) return fail(ctx, ZKMSM_ERR_CUDA, "fr_aggregate: %s", cudaGetErrorString(ex.err));
  CU(ctx, cudaMemcpyAsync(out, d_out, ob, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return ZKMSM_OK;
}

// quotient polynomial by transforms (fr_ntt.cuh): tables for n are built on first use and kept in the context
static int fr_quotient_ntt(zkmsm_ctx* ctx, const uint32_t* u, const uint32_t* v, const uint32_t* w, size_t n,
                           uint32_t* h_out, int* out_exact) {
  const FrNttPlan p = FrNttPlan::make((uint32_t)n);
  const size_t in_bytes = sizeof(uint32_t) * 8 * n, m = p.m;
  // WS_MISC: raw u | v | w | out (n x 8 words each), flag, then 5 m Fr of scratch
  size_t off_fr = (4 * in_bytes + 256 + 31) / 32 * 32;
  int rc = ws_reserve(ctx, WS_MISC, off_fr + sizeof(Fr) * 5 * m);
  if (rc) return rc;
  char* base = (char*)ctx->ws[WS_MISC];
  uint32_t* d_raw = (uint32_t*)base;
  uint32_t* d_out = d_raw + 24 * n;
  uint32_t* d_flag = d_raw + 32 * n;
  Fr* scratch = (Fr*)(base + off_fr);
  if (ctx->prof) ctx->prof->n = 0;
  CudaExec ex(ctx->stream, ctx->prof, ctx->tune.no_coop != 0, ctx->tune.ntt_no_fuse != 0);
  FrNttTables tb;
  if (ctx->ntt_n != n) {
    if (ctx->ntt_tables) { CU(ctx, cudaStreamSynchronize(ctx->stream)); cudaFree(ctx->ntt_tables); ctx->ntt_tables = nullptr; }
    ctx->ntt_n = 0;
    if (cudaMalloc(&ctx->ntt_tables, sizeof(Fr) * fr_ntt_table_elems(p)) != cudaSuccess) {
      cudaGetLastError();
      ctx->ntt_tables = nullptr;
      return fail(ctx, ZKMSM_ERR_NOMEM, "fr_quotient: out of device memory for the transform tables");
    }
    fr_ntt_tables_at(tb, p, ctx->ntt_tables);
    fr_quotient_setup(ex, p, tb, scratch);
    if (ex.err != cudaSuccess) return fail(ctx, ZKMSM_ERR_CUDA, "fr_quotient setup: %s", cudaGetErrorString(ex.err));
    ctx->ntt_n = n;
  }
  fr_ntt_tables_at(tb, p, ctx->ntt_tables);
  CU(ctx, cudaMemcpyAsync(d_raw, u, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
  CU(ctx, cudaMemcpyAsync(d_raw + 8 * n, v, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
  CU(ctx, cudaMemcpyAsync(d_raw + 16 * n, w, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
  CU(ctx, cudaMemsetAsync(d_flag, 0, sizeof(uint32_t), ctx->stream));
  fr_quotient_run(ex, p, tb, (const uint32_t*)d_raw, (const uint32_t*)(d_raw + 8 * n), (const uint32_t*)(d_raw + 16 * n), scratch,
                  d_out, d_flag);
  if (ex.err != cudaSuccess) return fail(ctx, ZKMSM_ERR_CUDA, "fr_quotient: %s", cudaGetErrorString(ex.err));
  ctx->last_launches = ex.launches;
  uint32_t flag = 0;
  CU(ctx, cudaMemcpyAsync(h_out, d_out, sizeof(uint32_t) * 8 * (n - 1), cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaMemcpyAsync(&flag, d_flag, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  *out_exact = flag ? 0 : 1;
  return ZKMSM_OK;
}

// quotient polynomial h = (u v - w) / t,  t = prod_{k=1..n} (x - k)
extern "C" int zkmsm_fr_quotient(zkmsm_ctx* ctx, const uint32_t* u, const uint32_t* v, const uint32_t* w, size_t n,
                                 uint32_t* h_out, int* out_exact) {
  if (!ctx || !u || !v || !w || !h_out || !out_exact || n < 2 || n > (1u << 22))
    return fail(ctx, ZKMSM_ERR_INVALID_ARG, "bad argument (2 <= n <= 2^22)");
  CU(ctx, cudaSetDevice(ctx->device));
  // transforms from 32 coefficients on; below that (and on request, as the cross-check) the reference's schoolbook
  // multiplication and long division, one launch per step
  const bool schoolbook = ctx->tune.quotient_schoolbook != 0;
  if (n >= 32 && !schoolbook) return fr_quotient_ntt(ctx, u, v, w, n, h_out, out_exact);
  if (n > (1u << 14)) return fail(ctx, ZKMSM_ERR_INVALID_ARG, "schoolbook quotient: n <= 2^14");
  const size_t in_bytes = sizeof(uint32_t) * 8 * n;
  // layout in WS_MISC: raw u | v | w, then Fr arrays u, v, w (n each), p (2n), t0, t1 (n+1 each), h (n), flag
  size_t off_raw = 0, off_fr = 3 * in_bytes;
  size_t fr_elems = 3 * n + 2 * n + 2 * (n + 1) + n;
  int rc = ws_reserve(ctx, WS_MISC, off_fr + sizeof(Fr) * fr_elems + in_bytes + 256);
  if (rc) return rc;
  char* base = (char*)ctx->ws[WS_MISC];
  uint32_t* d_raw = (uint32_t*)(base + off_raw);
  Fr* du = (Fr*)(base + off_fr);
  Fr *dv = du + n, *dw = dv + n, *dp = dw + n, *dt0 = dp + 2 * n, *dt1 = dt0 + (n + 1), *dh = dt1 + (n + 1);
  uint32_t* d_out = (uint32_t*)(dh + n);
  uint32_t* d_flag = d_out + 8 * n;
  CU(ctx, cudaMemcpyAsync(d_raw, u, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
  CU(ctx, cudaMemcpyAsync(d_raw + 8 * n, v, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
  CU(ctx, cudaMemcpyAsync(d_raw + 16 * n, w, in_bytes, cudaMemcpyHostToDevice, ctx->stream));
  CU(ctx, cudaMemsetAsync(d_flag, 0, sizeof(uint32_t), ctx->stream));
  CU(ctx, cudaMemsetAsync(dt0, 0, sizeof(Fr) * 2 * (n + 1), ctx->stream));
  CudaExec ex(ctx->stream);
  const uint32_t N = (uint32_t)n;
  ex.template launch<FrVecToMont>(N, N, N, (const uint32_t*)d_raw, du);
  ex.template launch<FrVecToMont>(N, N, N, (const uint32_t*)(d_raw + 8 * n), dv);
  ex.template launch<FrVecToMont>(N, N, N, (const uint32_t*)(d_raw + 16 * n), dw);
  ex.template launch<FrPolyMulSub>(2 * N - 1, N, (const Fr*)du, (const Fr*)dv, (const Fr*)dw, dp);
  // t_0 = 1 (Montgomery one), then n steps t_k = t_{k-1} (x - k)
  Fr one_m, k_m;
  for (int i = 0; i < 8; i++) one_m.v[i] = FR_ONE_HOST[i];
  CU(ctx, cudaMemcpyAsync(dt0, &one_m, sizeof(Fr), cudaMemcpyHostToDevice, ctx->stream));
  Fr* t_old = dt0;
  Fr* t_new = dt1;
  fset_zero(k_m);
  for (uint32_t k = 1; k <= N; k++) {
    // k in Montgomery form = k * R mod r, accumulated on the host by adding R mod r (exact 256-bit arithmetic)
    fr_host_add_one_mont(k_m);
    ex.template launch<FrTStep>(k + 1, k, k_m, (const Fr*)t_old, t_new);
    Fr* tmp = t_old; t_old = t_new; t_new = tmp;
  }
  // long division: quotient degree (2n-2) - n = n-2
  for (int d = (int)n - 2; d >= 0; d--) ex.template launch<FrDivStep>(N, N, (uint32_t)d, (const Fr*)t_old, dp, dh);
  ex.template launch<FrQuotientOut>(N, N, (const Fr*)dh, (const Fr*)dp, d_out, d_flag);
  if (ex.err != cudaSuccess) return fail(ctx, ZKMSM_ERR_CUDA, "fr_quotient: %s", cudaGetErrorString(ex.err));
  uint32_t flag = 0;
  CU(ctx, cudaMemcpyAsync(h_out, d_out, sizeof(uint32_t) * 8 * (n - 1), cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaMemcpyAsync(&flag, d_flag, sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  *out_exact = flag ? 0 : 1;
  return ZKMSM_OK;
}

// ------------------------------------------------------------------------------------------------
// Groth16: the resident CRS and Prover::prove as three concurrent MSMs (see Groth16Scalars in fr_ops.cuh)
struct zkmsm_crs {
  zkmsm_ctx* owner;
  zkmsm_ctx* lane[2];      // two more contexts on the owner's device: stream + workspace for the B and C MSMs
  zkmsm_points* set_A;     // g1.xi ++ [alpha, delta]
  zkmsm_points* set_B;     // g2.xi ++ [beta_2, delta_2]
  zkmsm_points* set_C;     // g1.uvw_wit ++ g1.xt_by_delta ++ g1.xi ++ [alpha, beta, delta]
  size_t n, n_wit, n_xt;
  uint32_t* d_scalars;     // su (n + 2) | sv (n + 2) | sc (n_wit + n_xt + n + 3) | r | s, 8 words each
  uint32_t* h_rs;          // pinned staging for r | s
  cudaEvent_t ready, done[2];
};

static void concat_points(uint32_t* dst, uint8_t* dinf, size_t& pos, int words, const uint32_t* src, const uint8_t* inf, size_t n) {
  if (n) memcpy(dst + pos * words, src, sizeof(uint32_t) * words * n);
  if (dinf) { if (inf) memcpy(dinf + pos, inf, n); else memset(dinf + pos, 0, n); }
  pos += n;
}

extern "C" int zkmsm_crs_free(zkmsm_ctx* ctx, zkmsm_crs* crs) {
  if (!crs) return ZKMSM_ERR_INVALID_ARG;
  zkmsm_ctx* o = crs->owner ? crs->owner : ctx;
  if (o) cudaSetDevice(o->device);
  for (int i = 0; i < 2; i++)
    if (crs->lane[i]) { cudaStreamSynchronize(crs->lane[i]->stream); }
  if (crs->set_A) zkmsm_points_free(o, crs->set_A);
  if (crs->set_B) { if (crs->lane[0]) graphs_drop(crs->lane[0], crs->set_B->uid); zkmsm_points_free(o, crs->set_B); }
  if (crs->set_C) { if (crs->lane[1]) graphs_drop(crs->lane[1], crs->set_C->uid); zkmsm_points_free(o, crs->set_C); }
  for (int i = 0; i < 2; i++)
    if (crs->lane[i]) zkmsm_destroy(crs->lane[i]);
  if (crs->d_scalars) cudaFree(crs->d_scalars);
  if (crs->h_rs) cudaFreeHost(crs->h_rs);
  if (crs->ready) cudaEventDestroy(crs->ready);
  for (int i = 0; i < 2; i++)
    if (crs->done[i]) cudaEventDestroy(crs->done[i]);
  delete crs;
  return ZKMSM_OK;
}

extern "C" int zkmsm_crs_load(zkmsm_ctx* ctx, const zkmsm_crs_desc* d, unsigned flags, zkmsm_crs** out) {
  if (!ctx || !d || !out) return fail(ctx, ZKMSM_ERR_INVALID_ARG, "null argument");
  *out = nullptr;
  if (!d->g1_alpha || !d->g1_beta || !d->g1_delta || !d->g2_beta || !d->g2_delta || (d->n && (!d->g1_xi || !d->g2_xi)) ||
      (d->n_wit && !d->g1_uvw_wit) || (d->n_xt && !d->g1_xt_by_delta))
    return fail(ctx, ZKMSM_ERR_INVALID_ARG, "CRS descriptor: null vector");
  if (d->n_xt > d->n) return fail(ctx, ZKMSM_ERR_INVALID_ARG, "CRS descriptor: xt_by_delta longer than xi");
  CU(ctx, cudaSetDevice(ctx->device));
  zkmsm_crs* crs = new (std::nothrow) zkmsm_crs();
  if (!crs) return fail(ctx, ZKMSM_ERR_NOMEM, "host allocation");
  memset(crs, 0, sizeof(*crs));
  crs->owner = ctx;
  crs->n = d->n; crs->n_wit = d->n_wit; crs->n_xt = d->n_xt;
  int rc = ZKMSM_OK;
  const size_t nA = d->n + 2, nC = d->n_wit + d->n_xt + d->n + 3;
  const bool any_inf = d->g1_xi_inf || d->g1_uvw_wit_inf || d->g1_xt_by_delta_inf || d->g2_xi_inf;
  uint32_t* buf = (uint32_t*)malloc(sizeof(uint32_t) * 48 * (nC > nA ? nC : nA));
  uint8_t* binf = any_inf ? (uint8_t*)malloc(nC > nA ? nC : nA) : nullptr;
  if (!buf || (any_inf && !binf)) rc = fail(ctx, ZKMSM_ERR_NOMEM, "host allocation");
  size_t pos;
  if (!rc) {
    pos = 0;
    concat_points(buf, binf, pos, 24, d->g1_xi, d->g1_xi_inf, d->n);
    concat_points(buf, binf, pos, 24, d->g1_alpha, nullptr, 1);
    concat_points(buf, binf, pos, 24, d->g1_delta, nullptr, 1);
    rc = load_points_impl<G1>(ctx, buf, binf, nA, flags, 1, &crs->set_A);
  }
  if (!rc) {
    pos = 0;
    concat_points(buf, binf, pos, 48, d->g2_xi, d->g2_xi_inf, d->n);
    concat_points(buf, binf, pos, 48, d->g2_beta, nullptr, 1);
    concat_points(buf, binf, pos, 48, d->g2_delta, nullptr, 1);
    rc = load_points_impl<G2>(ctx, buf, binf, nA, flags, 2, &crs->set_B);
  }
  if (!rc) {
    pos = 0;
    concat_points(buf, binf, pos, 24, d->g1_uvw_wit, d->g1_uvw_wit_inf, d->n_wit);
    concat_points(buf, binf, pos, 24, d->g1_xt_by_delta, d->g1_xt_by_delta_inf, d->n_xt);
    concat_points(buf, binf, pos, 24, d->g1_xi, d->g1_xi_inf, d->n);
    concat_points(buf, binf, pos, 24, d->g1_alpha, nullptr, 1);
    concat_points(buf, binf, pos, 24, d->g1_beta, nullptr, 1);
    concat_points(buf, binf, pos, 24, d->g1_delta, nullptr, 1);
    rc = load_points_impl<G1>(ctx, buf, binf, nC, flags, 1, &crs->set_C);
  }
  free(buf);
  free(binf);
  if (!rc && (zkmsm_create(ctx->device, &crs->lane[0]) || zkmsm_create(ctx->device, &crs->lane[1])))
    rc = fail(ctx, ZKMSM_ERR_CUDA, "could not create the prover's extra streams");
  if (!rc) {
    cudaSetDevice(ctx->device);
    const size_t words = 8 * (2 * nA + nC + 2);
    if (cudaMalloc(&crs->d_scalars, sizeof(uint32_t) * words) != cudaSuccess || cudaMallocHost(&crs->h_rs, 64) != cudaSuccess ||
        cudaEventCreateWithFlags(&crs->ready, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&crs->done[0], cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&crs->done[1], cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      rc = fail(ctx, ZKMSM_ERR_NOMEM, "prover scalar buffers");
    }
  }
  if (rc) { zkmsm_crs_free(ctx, crs); return rc; }
  for (int i = 0; i < 2; i++) { crs->lane[i]->tune = ctx->tune; crs->lane[i]->window_override = ctx->window_override; }
  // Stream priorities: the G2 MSM (lane 0) has the longest latency-bound tail, then C (lane 1); with the higher
  // priority their bucket accumulation goes first and their tails then run under the other MSMs' accumulation
  // instead of at the end of the proof.
  {
    int least = 0, greatest = 0;
    if (cudaDeviceGetStreamPriorityRange(&least, &greatest) == cudaSuccess && greatest < least) {
      for (int i = 0; i < 2; i++) {
        cudaStream_t st = nullptr;
        int prio = greatest + i < least ? greatest + i : least;
        if (cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, prio) == cudaSuccess) {
          cudaStreamDestroy(crs->lane[i]->own_stream);
          crs->lane[i]->own_stream = crs->lane[i]->stream = st;
        }
      }
    }
    cudaGetLastError();
  }
  *out = crs;
  return ZKMSM_OK;
}

extern "C" int zkmsm_crs_sizes(const zkmsm_crs* crs, size_t* n, size_t* n_wit, size_t* n_xt) {
  if (!crs) return ZKMSM_ERR_INVALID_ARG;
  if (n) *n = crs->n;
  if (n_wit) *n_wit = crs->n_wit;
  if (n_xt) *n_xt = crs->n_xt;
  return ZKMSM_OK;
}

// copies the four coefficient vectors into place, fills the derived scalars and enqueues the three MSMs on three
// streams; rank/world as in zkmsm_g1_msm_partial_range.  The results stay in the three contexts' result blocks.
static int groth16_enqueue(zkmsm_ctx* ctx, zkmsm_crs* crs, const uint32_t* u, const uint32_t* v, const uint32_t* h,
                           const uint32_t* wit, const uint32_t* r, const uint32_t* s, bool want_affine, uint32_t rank, uint32_t world) {
  if (!ctx || !crs || !r || !s || (crs->n && (!u || !v)) || (crs->n_xt && !h) || (crs->n_wit && !wit))
    return fail(ctx, ZKMSM_ERR_INVALID_ARG, "null argument");
  if (crs->owner != ctx) return fail(ctx, ZKMSM_ERR_INVALID_ARG, "the CRS belongs to another context");
  CU(ctx, cudaSetDevice(ctx->device));
  const size_t n = crs->n, nA = n + 2, nC = crs->n_wit + crs->n_xt + n + 3;
  uint32_t* su = crs->d_scalars;
  uint32_t* sv = su + 8 * nA;
  uint32_t* sc = sv + 8 * nA;
  uint32_t* d_rs = sc + 8 * nC;
  cudaStream_t st = ctx->stream;
  // the previous proof's B and C MSMs read sv / sc on their own streams: wait for them before overwriting
  CU(ctx, cudaStreamWaitEvent(st, crs->done[0], 0));
  CU(ctx, cudaStreamWaitEvent(st, crs->done[1], 0));
  memcpy(crs->h_rs, r, 32);
  memcpy(crs->h_rs + 8, s, 32);
  const size_t row = sizeof(uint32_t) * 8;
  zkmsm_ctx* cb = crs->lane[0];
  zkmsm_ctx* cc = crs->lane[1];
  CudaExec ex(st);
  // v first: B (the G2 MSM, the longest) reads only sv = v ++ [1, s] and starts while u, h and the witness are still
  // on their way (cudaMemcpyDefault: the four vectors may live in host memory or already on this device)
  CU(ctx, cudaMemcpyAsync(d_rs, crs->h_rs, 64, cudaMemcpyHostToDevice, st));
  if (n) CU(ctx, cudaMemcpyAsync(sv, v, row * n, cudaMemcpyDefault, st));
  ex.template launch<Groth16TailB>(1u, (uint32_t)n, (const uint32_t*)d_rs, sv);
  if (ex.err != cudaSuccess) return fail(ctx, ZKMSM_ERR_CUDA, "groth16 scalars: %s", cudaGetErrorString(ex.err));
  CU(ctx, cudaEventRecord(crs->ready, st));
  CU(ctx, cudaStreamWaitEvent(cb->stream, crs->ready, 0));
  int rc;
  if ((rc = msm_enqueue_impl<G2>(cb, crs->set_B, sv, nA, 2, want_affine, nullptr, rank, world))) { strcpy(ctx->err, cb->err); return rc; }
  CU(ctx, cudaEventRecord(crs->done[0], cb->stream));
  if (n) CU(ctx, cudaMemcpyAsync(su, u, row * n, cudaMemcpyDefault, st));
  if (crs->n_wit) CU(ctx, cudaMemcpyAsync(sc, wit, row * crs->n_wit, cudaMemcpyDefault, st));
  if (crs->n_xt) CU(ctx, cudaMemcpyAsync(sc + 8 * crs->n_wit, h, row * crs->n_xt, cudaMemcpyDefault, st));
  CU(ctx, cudaMemsetAsync(&ctx->d_res->aux_err, 0, sizeof(uint32_t), st));
  ex.template launch<Groth16Scalars>((uint32_t)n + 1, (uint32_t)n, (const uint32_t*)d_rs, su, sv, sc + 8 * (crs->n_wit + crs->n_xt), 0u,
                                     &ctx->d_res->aux_err);
  if (ex.err != cudaSuccess) return fail(ctx, ZKMSM_ERR_CUDA, "groth16 scalars: %s", cudaGetErrorString(ex.err));
  CU(ctx, cudaEventRecord(crs->ready, st));
  CU(ctx, cudaStreamWaitEvent(cc->stream, crs->ready, 0));
  if ((rc = msm_enqueue_impl<G1>(cc, crs->set_C, sc, nC, 1, want_affine, nullptr, rank, world))) { strcpy(ctx->err, cc->err); return rc; }
  CU(ctx, cudaEventRecord(crs->done[1], cc->stream));
  if ((rc = msm_enqueue_impl<G1>(ctx, crs->set_A, su, nA, 1, want_affine, nullptr, rank, world))) return rc;
  return ZKMSM_OK;
}

// waits for the three MSMs; the result blocks are then in the three contexts' pinned mirrors
static int groth16_collect(zkmsm_ctx* ctx, zkmsm_crs* crs) {
  int rc = msm_collect(ctx);
  int rb = msm_collect(crs->lane[0]), rcc = msm_collect(crs->lane[1]);
  if (!rc && rb) { strcpy(ctx->err, crs->lane[0]->err); rc = rb; }
  if (!rc && rcc) { strcpy(ctx->err, crs->lane[1]->err); rc = rcc; }
  if (!rc && ctx->h_res->aux_err) rc = fail(ctx, ZKMSM_ERR_SCALAR_RANGE, "r or s is not a field element (>= r)");
  return rc;
}

extern "C" int zkmsm_groth16_prove(zkmsm_ctx* ctx, zkmsm_crs* crs, const uint32_t* u, const uint32_t* v, const uint32_t* h,
                                   const uint32_t* wit, const uint32_t* r, const uint32_t* s, uint32_t* proof_out, int* inf_out) {
  if (!proof_out || !inf_out) return fail(ctx, ZKMSM_ERR_INVALID_ARG, "null output");
  int rc = groth16_enqueue(ctx, crs, u, v, h, wit, r, s, true, 0, 1);
  if (rc) return rc;
  if ((rc = groth16_collect(ctx, crs))) return rc;
  const ResultBlock* res[3] = {ctx->h_res, crs->lane[0]->h_res, crs->lane[1]->h_res};
  const int words[3] = {24, 48, 24}, at[3] = {0, 24, 72};
  for (int i = 0; i < 3; i++) {
    inf_out[i] = res[i]->inf ? 1 : 0;
    if (res[i]->inf) memset(proof_out + at[i], 0, sizeof(uint32_t) * words[i]);
    else memcpy(proof_out + at[i], res[i]->affine, sizeof(uint32_t) * words[i]);
  }
  return ZKMSM_OK;
}

extern "C" int zkmsm_groth16_prove_partial(zkmsm_ctx* ctx, zkmsm_crs* crs, const uint32_t* u, const uint32_t* v, const uint32_t* h,
                                           const uint32_t* wit, const uint32_t* r, const uint32_t* s, unsigned rank, unsigned world,
                                           uint32_t* out_partials) {
  if (!out_partials) return fail(ctx, ZKMSM_ERR_INVALID_ARG, "null output");
  int rc = groth16_enqueue(ctx, crs, u, v, h, wit, r, s, false, rank, world);
  if (rc) return rc;
  if ((rc = groth16_collect(ctx, crs))) return rc;
  memcpy(out_partials, ctx->h_res->xyzz, sizeof(uint32_t) * 48);
  memcpy(out_partials + 48, crs->lane[0]->h_res->xyzz, sizeof(uint32_t) * 96);
  memcpy(out_partials + 144, crs->lane[1]->h_res->xyzz, sizeof(uint32_t) * 48);
  return ZKMSM_OK;
}

extern "C" int zkmsm_groth16_combine(zkmsm_ctx* ctx, const uint32_t* partials, size_t world, uint32_t* proof_out, int* inf_out) {
  if (!ctx || !partials || !proof_out || !inf_out || world == 0 || world > 4096) return fail(ctx, ZKMSM_ERR_INVALID_ARG, "bad argument");
  CU(ctx, cudaSetDevice(ctx->device));
  // one upload of the per-rank blobs, three combine kernels reading them in place (stride = one blob), one read-back
  const size_t bytes = sizeof(uint32_t) * ZKMSM_GROTH16_PARTIAL_WORDS * world;
  int rc = ws_reserve(ctx, WS_MISC, bytes);
  if (rc) return rc;
  uint32_t* d = (uint32_t*)ctx->ws[WS_MISC];
  CU(ctx, cudaMemcpyAsync(d, partials, bytes, cudaMemcpyHostToDevice, ctx->stream));
  CU(ctx, cudaMemsetAsync(&ctx->d_res->err, 0, sizeof(uint32_t), ctx->stream));
  CudaExec ex(ctx->stream, nullptr, ctx->tune.no_coop != 0);
  ResultBlock* r = ctx->d_res;
  combine_launch<G1>(ctx, ex, (const XYZZ<Fp>*)d, (uint32_t)world, ZKMSM_GROTH16_PARTIAL_WORDS, r->proof, &r->proof_inf[0]);
  combine_launch<G2>(ctx, ex, (const XYZZ<Fp2>*)(d + 48), (uint32_t)world, ZKMSM_GROTH16_PARTIAL_WORDS, r->proof + 24, &r->proof_inf[1]);
  combine_launch<G1>(ctx, ex, (const XYZZ<Fp>*)(d + 144), (uint32_t)world, ZKMSM_GROTH16_PARTIAL_WORDS, r->proof + 72, &r->proof_inf[2]);
  if (ex.err != cudaSuccess) return fail(ctx, ZKMSM_ERR_CUDA, "groth16 combine: %s", cudaGetErrorString(ex.err));
  CU(ctx, cudaMemcpyAsync(ctx->h_res, ctx->d_res, sizeof(ResultBlock), cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  ctx->pending = 0;
  if (ctx->h_res->err & ERR_SCALAR_RANGE) return fail(ctx, ZKMSM_ERR_SCALAR_RANGE, "a rank's share saw a scalar out of range");
  const int words[3] = {24, 48, 24}, at[3] = {0, 24, 72};
  for (int i = 0; i < 3; i++) {
    inf_out[i] = ctx->h_res->proof_inf[i] ? 1 : 0;
    if (inf_out[i]) memset(proof_out + at[i], 0, sizeof(uint32_t) * words[i]);
    else memcpy(proof_out + at[i], ctx->h_res->proof + at[i], sizeof(uint32_t) * words[i]);
  }
  return ZKMSM_OK;
}

// ------------------------------------------------------------------------------------------------
// integer-multiply throughput probe (kernels in tu_probe.cu)
extern "C" int zkmsm_probe_launch(int variant, int grid, int block, cudaStream_t st, uint32_t* sink, int iters);

extern "C" int zkmsm_bench_imad(zkmsm_ctx* ctx, int variant, int iters, double* out_lp_per_s, double* out_ms) {
  if (!ctx || !out_lp_per_s || iters <= 0 || variant < 0 || variant > 3) return fail(ctx, ZKMSM_ERR_INVALID_ARG, "bad argument");
  CU(ctx, cudaSetDevice(ctx->device));
  int sms = 0;
  CU(ctx, cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, ctx->device));
  int rc = ws_reserve(ctx, WS_MISC, 64);
  if (rc) return rc;
  dim3 grid(sms * 8), block(256);
  const double per_iter[4] = {32.0, 24.0, 32.0, 24.0};
  cudaEvent_t e0, e1;
  CU(ctx, cudaEventCreate(&e0));
  CU(ctx, cudaEventCreate(&e1));
  for (int rep = 0; rep < 2; rep++) {  // first pass warms up
    CU(ctx, cudaEventRecord(e0, ctx->stream));
    if (zkmsm_probe_launch(variant, (int)grid.x, (int)block.x, ctx->stream, (uint32_t*)ctx->ws[WS_MISC], iters))
      return fail(ctx, ZKMSM_ERR_CUDA, "probe launch failed");
    CU(ctx, cudaEventRecord(e1, ctx->stream));
    CU(ctx, cudaEventSynchronize(e1));
  }
  CU(ctx, cudaGetLastError());
  float ms = 0;
  CU(ctx, cudaEventElapsedTime(&ms, e0, e1));
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  double ops = (double)grid.x * block.x * (double)iters * per_iter[variant];
  *out_lp_per_s = ops / (ms * 1e-3);
  if (out_ms) *out_ms = ms;
  return ZKMSM_OK;
}
