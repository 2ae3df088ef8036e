// Two contexts in ONE process -- on two devices when the box has them -- through the C ABI only
// (include/zkmsm.h): point-split partials, bucket-range-split partials, combine; every result compared with the
// closed form (sum s_i k_i) g computed by the fixed-base kernel.  Covers the per-device shared-memory opt-in
// (zkmsm_create) that a single-process multi-GPU host depends on.
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>
#include "zkmsm.h"

static int failures = 0;
#define CHECK(cond)                                                                   \
  do {                                                                                \
    if (!(cond)) { printf("FAIL %s:%d  %s\n", __FILE__, __LINE__, #cond); failures++; } \
  } while (0)
#define OK(call)                                                                                        \
  do {                                                                                                  \
    int rc__ = (call);                                                                                  \
    if (rc__ != ZKMSM_OK) { printf("FAIL %s:%d  %s -> %d\n", __FILE__, __LINE__, #call, rc__); failures++; } \
  } while (0)

// generator of G1 (g1_point.rs:41-44), canonical little-endian limbs x | y
static const uint32_t G1_GEN[24] = {
    0xdb22c6bbu, 0xfb3af00au, 0xf97a1aefu, 0x6c55e83fu, 0x171bac58u, 0xa14e3a3fu, 0x9774b905u, 0xc3688c4fu, 0x4fa9ac0fu, 0x2695638cu,
    0x3197d794u, 0x17f1d3a7u,
    0x46c5e7e1u, 0x0caa2329u, 0xa2888ae4u, 0xd03cc744u, 0x2c04b3edu, 0x00db18cbu, 0xd5d00af6u, 0xfcf5e095u, 0x741d8ae4u, 0xa09e30edu,
    0xe3aaa0f1u, 0x08b3f481u};

int main() {
  zkmsm_ctx *c0 = nullptr, *c1 = nullptr;
  if (zkmsm_create(0, &c0) != ZKMSM_OK) { printf("no sm_100 device\n"); return 2; }
  int dev1 = zkmsm_create(1, &c1) == ZKMSM_OK ? 1 : 0;
  if (!dev1) OK(zkmsm_create(0, &c1));
  printf("second context on device %d\n", dev1);
  const size_t n = 6000;
  std::vector<uint32_t> k(n * 8, 0), s(n * 8, 0);
  unsigned __int128 total = 0;
  for (size_t i = 0; i < n; i++) {
    k[i * 8] = (uint32_t)(i + 1);
    uint64_t si = i * 7919u + 3;
    s[i * 8] = (uint32_t)si;
    s[i * 8 + 1] = (uint32_t)(si >> 32);
    total += (unsigned __int128)si * (i + 1);
  }
  uint32_t tot[8] = {0};
  for (int j = 0; j < 4; j++) tot[j] = (uint32_t)(total >> (32 * j));
  uint32_t want[24];
  uint8_t want_inf = 0;
  OK(zkmsm_g1_mul_base(c0, G1_GEN, tot, 1, want, &want_inf));
  CHECK(!want_inf);
  const unsigned flags = ZKMSM_PRECOMPUTE | ZKMSM_SUBGROUP;
  zkmsm_points *full0 = nullptr, *full1 = nullptr, *lo = nullptr, *hi = nullptr;
  OK(zkmsm_g1_points_from_scalars(c0, G1_GEN, k.data(), n, flags, &full0));
  OK(zkmsm_g1_points_from_scalars(c1, G1_GEN, k.data(), n, flags, &full1));
  OK(zkmsm_g1_points_from_scalars(c0, G1_GEN, k.data(), n / 2, flags, &lo));
  OK(zkmsm_g1_points_from_scalars(c1, G1_GEN, k.data() + 8 * (n / 2), n - n / 2, flags, &hi));
  uint32_t got[24], parts[2 * ZKMSM_G1_PARTIAL_WORDS];
  int inf = 0;
  // whole MSM on each device
  OK(zkmsm_g1_msm(c0, full0, s.data(), n, got, &inf));
  CHECK(!inf && memcmp(got, want, sizeof(want)) == 0);
  OK(zkmsm_g1_msm(c1, full1, s.data(), n, got, &inf));
  CHECK(!inf && memcmp(got, want, sizeof(want)) == 0);
  // point split: one contiguous shard per device
  OK(zkmsm_g1_msm_partial(c0, lo, s.data(), n / 2, parts));
  OK(zkmsm_g1_msm_partial(c1, hi, s.data() + 8 * (n / 2), n - n / 2, parts + ZKMSM_G1_PARTIAL_WORDS));
  OK(zkmsm_g1_combine(c0, parts, 2, got, &inf));
  CHECK(!inf && memcmp(got, want, sizeof(want)) == 0);
  // bucket-range split: both devices hold the whole set and see all scalars
  OK(zkmsm_g1_msm_partial_range(c0, full0, s.data(), n, 0, 2, parts));
  OK(zkmsm_g1_msm_partial_range(c1, full1, s.data(), n, 1, 2, parts + ZKMSM_G1_PARTIAL_WORDS));
  OK(zkmsm_g1_combine(c1, parts, 2, got, &inf));
  CHECK(!inf && memcmp(got, want, sizeof(want)) == 0);
  // a point set cannot be used from the other device's context
  if (dev1) CHECK(zkmsm_g1_msm(c1, full0, s.data(), n, got, &inf) == ZKMSM_ERR_INVALID_ARG);
  zkmsm_points_free(c0, full0);
  zkmsm_points_free(c1, full1);
  zkmsm_points_free(c0, lo);
  zkmsm_points_free(c1, hi);
  zkmsm_destroy(c0);
  zkmsm_destroy(c1);
  printf("OK: %d failure(s)\n", failures);
  return failures ? 1 : 0;
}
