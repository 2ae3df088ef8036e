/* zkmsm.h -- C ABI of the B200-native MSM / Groth16-prove path for zk-toolkit's BLS12-381.
 *
 * The reference (exfinen/zk-toolkit, Rust) has no FFI for its own curve code; its seams are
 * ordinary methods.  Each entry point below names the reference method it replaces
 * (paths relative to the reference's src/).  The call style -- one-time create, opaque
 * handles, out-parameters, integer status -- follows the reference's only native binding,
 * the mcl backend (building_block/mcl/mcl_initializer.rs:8-15, mcl_g1.rs:64-86).
 *
 * Data layout at this boundary (all host memory unless a name says `_device`):
 *   Fq element : 12 x uint32, little-endian limbs, canonical value in [0, q)
 *   Fr scalar  :  8 x uint32, little-endian limbs, value < 2^255 (reference scalars are Fr
 *                 elements, i.e. < r; prime_field_elem.rs:263-272 reduces on construction)
 *   G1 point   : 24 x uint32 = x | y                  (g1_point.rs:33-36  Rational{x,y})
 *   G2 point   : 48 x uint32 = x.u0 | x.u1 | y.u0 | y.u1   (g2_point.rs:31-34; NB the
 *                 reference's Fq2::new takes (u1, u0), fq2.rs:22)
 *   AtInfinity : a separate flag byte per point (nullable array = no point is infinity);
 *                results return it in *out_is_inf and zero the coordinates.
 *
 * Semantics kept from the reference: n = 0 gives AtInfinity
 * (polynomial.rs:276); only the first n loaded points are used (polynomial.rs:277); n larger
 * than the point set is an error where the reference panics (polynomial.rs:278); scalar 0 or
 * point AtInfinity contributes nothing (macros.rs:11,44-52); P+P and P+(-P) inside the sum are
 * exact (macros.rs:53-108); the result is the canonical affine point (g1_point.rs:163-174).
 *
 * Threading: one caller at a time per context (the reference is single-threaded).
 * There is NO CPU fallback: every call fails with ZKMSM_ERR_NO_DEVICE / ZKMSM_ERR_CUDA when no
 * sm_100 device is usable.
 */
#ifndef ZKMSM_H
#define ZKMSM_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ZKMSM_OK 0
#define ZKMSM_ERR_INVALID_ARG (-1)
#define ZKMSM_ERR_CUDA (-2)
#define ZKMSM_ERR_SCALAR_RANGE (-3)   /* a scalar had bit 255 set, or was >= r for a ZKMSM_SUBGROUP set */
#define ZKMSM_ERR_NO_DEVICE (-4)
#define ZKMSM_ERR_TOO_FEW_POINTS (-5) /* n > points loaded; the reference panics, polynomial.rs:278 */
#define ZKMSM_ERR_NOMEM (-6)
#define ZKMSM_ERR_NOT_IN_SUBGROUP (-7) /* ZKMSM_CHECK_SUBGROUP: a loaded point does not have order r */

/* zkmsm_*_load_points flags */
#define ZKMSM_PRECOMPUTE 1u /* also store 2^(c w) P for every window w (CRS-style static sets) */
#define ZKMSM_SUBGROUP 2u   /* caller asserts every point has order r (true for any CRS vector: multiples of the
                             * generator).  Scalars must then be < r; s > (r-1)/2 is evaluated as (r-s)(-P), which
                             * saves the carry window.  Without the flag scalars are raw integers < 2^255 and nothing
                             * is assumed about the points (curves/macros.rs:10-21). */

#define ZKMSM_CHECK_SUBGROUP 4u /* verify r P = AtInfinity for every point at load (one 255-bit multiplication per point
                                * on the device); a point outside the subgroup fails the load with
                                * ZKMSM_ERR_NOT_IN_SUBGROUP.  Without it ZKMSM_SUBGROUP is the caller's unchecked claim. */

typedef struct zkmsm_ctx zkmsm_ctx;       /* one CUDA device + stream + workspace */
typedef struct zkmsm_points zkmsm_points; /* device-resident point set (G1 or G2) */

const char* zkmsm_version(void);

/* One-time init for one device, like MclInitializer::init (mcl_initializer.rs:8-15). */
int zkmsm_create(int device, zkmsm_ctx** out);
int zkmsm_destroy(zkmsm_ctx* ctx);
/* Run on the caller's CUDA stream (cudaStream_t) instead of the context's own; NULL restores it. */
int zkmsm_set_stream(zkmsm_ctx* ctx, void* cuda_stream);
/* Text of the last error on this context (never NULL). */
const char* zkmsm_last_error(const zkmsm_ctx* ctx);
/* Override the window width c for later MSMs / precomputed loads (0 = automatic). */
int zkmsm_set_window(zkmsm_ctx* ctx, unsigned c);

/* Tuning / cross-check switches of this context.  The environment variables ZKMSM_<NAME> (upper case) are read ONCE,
 * in zkmsm_create; this call overrides one of them afterwards.  Names: batch_rounds (-1 = automatic), batch_T,
 * batch_g2, batch_blocks, L, K, no_wave_L, no_coop, ntt_no_fuse, quotient_schoolbook, no_graph, no_bucket_acc
 * ("Tuning and cross-check switches"). */
int zkmsm_set_option(zkmsm_ctx* ctx, const char* name, long value);

/* Pinned host memory for scalars / points (optional; any host pointer is accepted). */
int zkmsm_host_alloc(size_t bytes, void** out);
int zkmsm_host_free(void* p);

/* ---- point sets: the `powers: &[G1Point]` / `&[G2Point]` argument of
 * Polynomial::eval_with_g{1,2}_hidings (field/polynomial.rs:272-293), i.e. the CRS vectors
 * crs.g1.xi, crs.g2.xi, crs.g1.xt_by_delta, crs.g1.uvw_wit (groth16/zktoolkit_based/crs.rs:17-34),
 * uploaded once and kept resident. */
int zkmsm_g1_load_points(zkmsm_ctx* ctx, const uint32_t* xy, const uint8_t* inf_flags, size_t n,
                         unsigned flags, zkmsm_points** out);
int zkmsm_g2_load_points(zkmsm_ctx* ctx, const uint32_t* xy, const uint8_t* inf_flags, size_t n,
                         unsigned flags, zkmsm_points** out);
int zkmsm_points_free(zkmsm_ctx* ctx, zkmsm_points* pts);
size_t zkmsm_points_len(const zkmsm_points* pts);
/* Layout of a set: window width c and number of slabs (0 / 0 for a plain set, whose c is chosen per call),
 * and the two load flags.  Any out-pointer may be NULL. */
int zkmsm_points_info(const zkmsm_points* pts, unsigned* c, unsigned* windows, int* precomputed, int* subgroup);
/* Copy a point set back as canonical affine limbs (+ flags, nullable). */
int zkmsm_points_read(zkmsm_ctx* ctx, const zkmsm_points* pts, size_t first, size_t n, uint32_t* xy,
                      uint8_t* inf_flags);

/* ---- the MSM: Polynomial::eval_with_g1_hidings (polynomial.rs:272-281) and
 * eval_with_g2_hidings (:284-293); also the inline loop `sum += &crs.g1.uvw_wit[i] * ai`
 * (groth16/zktoolkit_based/prover.rs:128-131).  result = sum_{i<n} scalars[i] * points[i]. */
int zkmsm_g1_msm(zkmsm_ctx* ctx, const zkmsm_points* pts, const uint32_t* scalars, size_t n,
                 uint32_t out_xy[24], int* out_is_inf);
int zkmsm_g2_msm(zkmsm_ctx* ctx, const zkmsm_points* pts, const uint32_t* scalars, size_t n,
                 uint32_t out_xy[48], int* out_is_inf);
/* Same with scalars already in device memory (n x 8 uint32). */
int zkmsm_g1_msm_device(zkmsm_ctx* ctx, const zkmsm_points* pts, const uint32_t* scalars_device, size_t n,
                        uint32_t out_xy[24], int* out_is_inf);
int zkmsm_g2_msm_device(zkmsm_ctx* ctx, const zkmsm_points* pts, const uint32_t* scalars_device, size_t n,
                        uint32_t out_xy[48], int* out_is_inf);
/* One call with everything in host memory: upload points, multiply, free
 * (exactly the (coeffs, powers) -> point shape of polynomial.rs:272-293). */
int zkmsm_g1_msm_oneshot(zkmsm_ctx* ctx, const uint32_t* xy, const uint8_t* inf_flags, const uint32_t* scalars,
                         size_t n, uint32_t out_xy[24], int* out_is_inf);
int zkmsm_g2_msm_oneshot(zkmsm_ctx* ctx, const uint32_t* xy, const uint8_t* inf_flags, const uint32_t* scalars,
                         size_t n, uint32_t out_xy[48], int* out_is_inf);

/* Stream-ordered halves of the above, for pipelining and device-side timing:
 * enqueue all kernels (no host synchronisation), later fetch the result. */
int zkmsm_g1_msm_enqueue(zkmsm_ctx* ctx, const zkmsm_points* pts, const uint32_t* scalars_device, size_t n);
int zkmsm_g2_msm_enqueue(zkmsm_ctx* ctx, const zkmsm_points* pts, const uint32_t* scalars_device, size_t n);
int zkmsm_g1_msm_result(zkmsm_ctx* ctx, uint32_t out_xy[24], int* out_is_inf);
int zkmsm_g2_msm_result(zkmsm_ctx* ctx, uint32_t out_xy[48], int* out_is_inf);
/* Host-scalar variant of enqueue: copies the scalars (asynchronously when they live in pinned memory, see
 * zkmsm_host_alloc) and enqueues; zkmsm_g{1,2}_msm_result collects.  Several contexts on one device give
 * concurrent MSMs (the five of a Groth16 proof, prover.rs:108-133, are independent). */
int zkmsm_g1_msm_begin(zkmsm_ctx* ctx, const zkmsm_points* pts, const uint32_t* scalars, size_t n);
int zkmsm_g2_msm_begin(zkmsm_ctx* ctx, const zkmsm_points* pts, const uint32_t* scalars, size_t n);
/* Number of kernels the last enqueue launched. */
int zkmsm_last_launch_count(const zkmsm_ctx* ctx);
/* Per-kernel timing of the MSMs that follow: CUDA events on the launching stream around every
 * launch.  zkmsm_profile_read waits for the stream and returns the number of launches of the last
 * MSM, filling (nullable) names[i * name_stride], ms[i] and threads[i]. */
int zkmsm_profile(zkmsm_ctx* ctx, int enable);
int zkmsm_profile_read(zkmsm_ctx* ctx, int max_entries, char* names, size_t name_stride, float* ms,
                       uint32_t* threads);

/* ---- multi-GPU: each rank multiplies its shard and exports one partial point (opaque
 * little-endian blob: 48 words for G1, 96 for G2); any rank adds the gathered partials.
 * Replaces the running `sum` of polynomial.rs:276-280 across shards. */
#define ZKMSM_G1_PARTIAL_WORDS 48
#define ZKMSM_G2_PARTIAL_WORDS 96
int zkmsm_g1_msm_partial(zkmsm_ctx* ctx, const zkmsm_points* pts, const uint32_t* scalars, size_t n,
                         uint32_t out_partial[ZKMSM_G1_PARTIAL_WORDS]);
int zkmsm_g2_msm_partial(zkmsm_ctx* ctx, const zkmsm_points* pts, const uint32_t* scalars, size_t n,
                         uint32_t out_partial[ZKMSM_G2_PARTIAL_WORDS]);
/* Stream-ordered variants (scalars and the partial in device memory, no host synchronisation).  An out-of-range
 * scalar cannot be reported by a call that does not wait: the partial is then exported poisoned and the combine
 * call that meets it returns ZKMSM_ERR_SCALAR_RANGE. */
int zkmsm_g1_msm_partial_device(zkmsm_ctx* ctx, const zkmsm_points* pts, const uint32_t* scalars_device, size_t n,
                                uint32_t* out_partial_device);
int zkmsm_g2_msm_partial_device(zkmsm_ctx* ctx, const zkmsm_points* pts, const uint32_t* scalars_device, size_t n,
                                uint32_t* out_partial_device);
/* The other way to split one MSM over `world` devices (a power of two): every device holds the WHOLE precomputed
 * point set and sees all n scalars, but owns only 1/world of the buckets (stripes of consecutive buckets dealt out
 * round-robin, so that every rank sees the same load), so that sorted pairs AND buckets (the latency-bound
 * reduction) shrink by `world`.  The partials of ranks 0..world-1 add up to the MSM exactly as
 * above.  Needs a ZKMSM_PRECOMPUTE set; world = 1 is the plain partial. */
int zkmsm_g1_msm_partial_range(zkmsm_ctx* ctx, const zkmsm_points* pts, const uint32_t* scalars, size_t n, unsigned rank,
                               unsigned world, uint32_t out_partial[ZKMSM_G1_PARTIAL_WORDS]);
int zkmsm_g2_msm_partial_range(zkmsm_ctx* ctx, const zkmsm_points* pts, const uint32_t* scalars, size_t n, unsigned rank,
                               unsigned world, uint32_t out_partial[ZKMSM_G2_PARTIAL_WORDS]);
int zkmsm_g1_msm_partial_range_device(zkmsm_ctx* ctx, const zkmsm_points* pts, const uint32_t* scalars_device, size_t n,
                                      unsigned rank, unsigned world, uint32_t* out_partial_device);
int zkmsm_g2_msm_partial_range_device(zkmsm_ctx* ctx, const zkmsm_points* pts, const uint32_t* scalars_device, size_t n,
                                      unsigned rank, unsigned world, uint32_t* out_partial_device);
int zkmsm_g1_combine(zkmsm_ctx* ctx, const uint32_t* partials, size_t k, uint32_t out_xy[24], int* out_is_inf);
int zkmsm_g2_combine(zkmsm_ctx* ctx, const uint32_t* partials, size_t k, uint32_t out_xy[48], int* out_is_inf);
int zkmsm_g1_combine_device(zkmsm_ctx* ctx, const uint32_t* partials_device, size_t k, uint32_t out_xy[24],
                            int* out_is_inf);
int zkmsm_g2_combine_device(zkmsm_ctx* ctx, const uint32_t* partials_device, size_t k, uint32_t out_xy[48],
                            int* out_is_inf);
/* Stream-ordered form: enqueue the sum of k device-resident partials; zkmsm_g{1,2}_msm_result fetches it. */
int zkmsm_g1_combine_enqueue(zkmsm_ctx* ctx, const uint32_t* partials_device, size_t k);
int zkmsm_g2_combine_enqueue(zkmsm_ctx* ctx, const uint32_t* partials_device, size_t k);

/* ---- vector scalar multiplication of one base point: `&G1Point * &Fq1`
 * (impl_scalar_mul_point!, curves/macros.rs:2-32) for n scalars at once, as CRS::new does
 * (crs.rs:88-116).  Scalars are full 256-bit raw integers, not reduced (macros.rs:10-21). */
int zkmsm_g1_mul_base(zkmsm_ctx* ctx, const uint32_t base_xy[24], const uint32_t* scalars, size_t n,
                      uint32_t* out_xy, uint8_t* out_inf_flags);
int zkmsm_g2_mul_base(zkmsm_ctx* ctx, const uint32_t base_xy[48], const uint32_t* scalars, size_t n,
                      uint32_t* out_xy, uint8_t* out_inf_flags);
/* Same, leaving the products on the device as a point set. */
int zkmsm_g1_points_from_scalars(zkmsm_ctx* ctx, const uint32_t base_xy[24], const uint32_t* scalars, size_t n,
                                 unsigned flags, zkmsm_points** out);
int zkmsm_g2_points_from_scalars(zkmsm_ctx* ctx, const uint32_t base_xy[48], const uint32_t* scalars, size_t n,
                                 unsigned flags, zkmsm_points** out);

/* ---- witness aggregation in Fr (the step in front of the MSMs): out[j] = sum_i wires[i] * polys[i][j] mod r,
 * i.e. the per-wire loop of Prover::prove (groth16/zktoolkit_based/prover.rs:108-117) / QAP::build_p's
 * `v += &self.vi[i] * wit` (qap/qap.rs:99-109) collapsed into one coefficient vector.  polys: n_wires x n
 * elements (8 words each, canonical, row i = wire i, zero padded), wires: n_wires elements, out: n elements. */
int zkmsm_fr_aggregate(zkmsm_ctx* ctx, const uint32_t* polys, size_t n_wires, size_t n, const uint32_t* wires,
                       uint32_t* out);

/* ---- quotient polynomial of the prover: h = (u v - w) / t with t = prod_{k=1..n} (x - k), i.e. Prover::new's
 * `p.divide_by(&t)` (groth16/zktoolkit_based/prover.rs:64-71) over QAP::build_p (qap/qap.rs:99-112) and
 * QAP::build_t (qap.rs:115-135); schoolbook product and long division as the reference (polynomial.rs:173-238).
 * u, v, w: n coefficients each (8 words, canonical, zero padded; coefficient i multiplies x^i); h_out: n-1
 * coefficients; *out_exact = 0 when the division leaves a remainder (the reference panics "p should be
 * divisible by t").  2 <= n <= 2^22.  From 32 coefficients on the division runs on number-theoretic transforms
 * (O(n log n); the tables for n are built on the first call and kept in the context); the coefficients are those of
 * the reference's schoolbook product and long division, which remain the path below 32 and the cross-check. */
int zkmsm_fr_quotient(zkmsm_ctx* ctx, const uint32_t* u, const uint32_t* v, const uint32_t* w, size_t n,
                      uint32_t* h_out, int* out_exact);

/* ---- Groth16 proof generation: Prover::prove (groth16/zktoolkit_based/prover.rs:96-147) in one call.
 * zkmsm_crs_load uploads the CRS of crs.rs:17-43 once (what the prover reads of it) and keeps it resident;
 * zkmsm_groth16_prove takes the prover's aggregated coefficient vectors
 *   u_j = sum_i a_i u_{i,j},  v_j = sum_i a_i v_{i,j}   (the per-wire loop of prover.rs:108-117 collapsed,
 *                                                         zkmsm_fr_aggregate),
 *   h = (u v - w) / t  (prover.rs:64-71, zkmsm_fr_quotient), the witness wires a_{l+1..m}, and the blinding
 * scalars r, s (the reference draws them itself, prover.rs:100-101; here the caller does, so that proofs are
 * reproducible), and returns Proof{A, B, C} (proof.rs:7-11) as canonical limbs A (24) | B (48) | C (24) with one
 * AtInfinity flag each.  A, B and C are three MSMs running concurrently on three streams; C uses the identity
 *   s A + r B_g1 - r s delta = sum_j (s u_j + r v_j) [x^j]_1 + s alpha + r beta + r s delta
 * so that B_g1 (prover.rs:120) and the two scalar multiplications of prover.rs:137-138 are folded into its MSM.
 * All vectors: 8 words per element, canonical, < r, in host memory or already on the context's device (a
 * multi-GPU host can upload one slice per device and all-gather them over NVLink).  flags as for zkmsm_*_load_points (a CRS is made of multiples
 * of the generators: ZKMSM_PRECOMPUTE | ZKMSM_SUBGROUP is the intended use).  `inf` arrays are optional
 * AtInfinity flags (NULL = none).  n_xt <= n. */
typedef struct zkmsm_crs zkmsm_crs;
typedef struct zkmsm_crs_desc {
  const uint32_t* g1_alpha;        /* crs.g1.alpha, 24 words */
  const uint32_t* g1_beta;         /* crs.g1.beta */
  const uint32_t* g1_delta;        /* crs.g1.delta */
  const uint32_t* g1_xi;           /* crs.g1.xi, n x 24 (crs.rs:88-102) */
  const uint8_t* g1_xi_inf;
  size_t n;
  const uint32_t* g1_uvw_wit;      /* crs.g1.uvw_wit, n_wit x 24 (crs.rs:65-86) */
  const uint8_t* g1_uvw_wit_inf;
  size_t n_wit;
  const uint32_t* g1_xt_by_delta;  /* crs.g1.xt_by_delta, n_xt x 24 (crs.rs:104-116) */
  const uint8_t* g1_xt_by_delta_inf;
  size_t n_xt;
  const uint32_t* g2_beta;         /* crs.g2.beta, 48 words */
  const uint32_t* g2_delta;        /* crs.g2.delta */
  const uint32_t* g2_xi;           /* crs.g2.xi, n x 48 */
  const uint8_t* g2_xi_inf;
} zkmsm_crs_desc;
int zkmsm_crs_load(zkmsm_ctx* ctx, const zkmsm_crs_desc* desc, unsigned flags, zkmsm_crs** out);
int zkmsm_crs_free(zkmsm_ctx* ctx, zkmsm_crs* crs);
int zkmsm_crs_sizes(const zkmsm_crs* crs, size_t* n, size_t* n_wit, size_t* n_xt);
int zkmsm_groth16_prove(zkmsm_ctx* ctx, zkmsm_crs* crs, const uint32_t* u, const uint32_t* v, const uint32_t* h,
                        const uint32_t* witness, const uint32_t r[8], const uint32_t s[8], uint32_t proof_out[96],
                        int out_is_inf[3]);
/* The same proof over `world` devices (configs[4]): every device loads the CRS and runs its 1/world share of the
 * three MSMs (bucket-range split, zkmsm_g1_msm_partial_range); the shares (one blob of
 * ZKMSM_GROTH16_PARTIAL_WORDS per rank, rank order, gathered by the caller -- NCCL, MPI or memcpy) are added by
 * zkmsm_groth16_combine on any one device.  Bit-identical to zkmsm_groth16_prove. */
#define ZKMSM_GROTH16_PARTIAL_WORDS 192 /* A (48) | B (96) | C (48) */
int zkmsm_groth16_prove_partial(zkmsm_ctx* ctx, zkmsm_crs* crs, const uint32_t* u, const uint32_t* v, const uint32_t* h,
                                const uint32_t* witness, const uint32_t r[8], const uint32_t s[8], unsigned rank,
                                unsigned world, uint32_t out_partials[ZKMSM_GROTH16_PARTIAL_WORDS]);
int zkmsm_groth16_combine(zkmsm_ctx* ctx, const uint32_t* partials, size_t world, uint32_t proof_out[96],
                          int out_is_inf[3]);

/* ---- diagnostics: integer-multiply throughput of this device (roofline denominator).
 * variant 0: independent mad.wide.u32; 1: carry-chained IMAD.WIDE.U32.X (mad.lo.cc/madc.hi.cc
 * pairs); 2: 32-bit IMAD (half a limb product each); 3: as 1 with data-dependent multipliers (the
 * Montgomery inner loop's shape).  Returns multiply-accumulates per second. */
int zkmsm_bench_imad(zkmsm_ctx* ctx, int variant, int iters, double* out_lp_per_s, double* out_ms);

#ifdef __cplusplus
}
#endif
#endif /* ZKMSM_H */
