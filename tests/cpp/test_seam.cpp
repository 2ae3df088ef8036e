// C++ parity test of the host-side mirror (include/zkmsm.hpp), written after the reference's own tests:
//   g1_point.rs:203-296 (scalar_mul, add_same_point, negate, vertical line, infinity cases),
//   g1_point.rs:347-371 (multiples of g), g2_point.rs:178-289, polynomial.rs:1250-1285 (eval_with_g1_hidings).
// Golden numbers come from tests/golden/ref_g{1,2}.json, handed over as "name value..." lines by the pytest
// wrapper (tests/test_gpu_cpp_host.py).  Exit code 0 = all checks passed.
#include <cstdio>
#include <fstream>
#include <iostream>
#include <map>
#include <sstream>

#include "zkmsm.hpp"

using namespace zk_toolkit;

static int failures = 0;
#define CHECK(cond)                                                       \
  do {                                                                    \
    if (!(cond)) { std::printf("FAILED %s:%d  %s\n", __FILE__, __LINE__, #cond); failures++; } \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 2) { std::printf("usage: test_seam kats.txt\n"); return 2; }
  std::map<std::string, std::vector<std::string>> kat;
  std::ifstream in(argv[1]);
  for (std::string line; std::getline(in, line);) {
    std::istringstream ss(line);
    std::string name, v;
    ss >> name;
    while (ss >> v) kat[name].push_back(v);
  }
  auto s = [](uint64_t v) { return Scalar::from_u64(v); };

  // ---- G1 (g1_point.rs)
  G1Point g = G1Point::g(), inf = G1Point::zero();
  CHECK(g * s(1) == g);                                       // scalar_mul :203-221
  CHECK(g * s(2) == g + g);
  CHECK(g * s(3) == g + g + g);
  {
    auto& k = kat["g1_add_same_point"];                       // add_same_point :223-237
    CHECK(g + g == G1Point::new_(Fq1::parse(k[0]), Fq1::parse(k[1])));
  }
  CHECK((g + (-g)).is_zero());                                // negate / add_vertical_line :239-262
  CHECK(g + inf == g);                                        // add_inf_and_affine :264-286
  CHECK(inf + g == g);
  CHECK((inf + inf).is_zero());                               // add_inf_and_inf :288-296
  std::vector<G1Point> gs;
  for (int n = 1; n <= 10; n++) {                             // scalar_mul_smaller_nums :347-356
    auto& k = kat["g1_mult_" + std::to_string(n)];
    gs.push_back(G1Point::new_(Fq1::parse(k[0]), Fq1::parse(k[1])));
    CHECK(g * s(n) == gs.back());
  }
  CHECK(gs[2] + gs[3] == gs[6]);                              // add_different_points :389-412
  CHECK(gs[8] + gs[0] == gs[9]);
  for (auto& kv : kat)                                        // scalar_mul_gen_pubkey :352-371
    if (kv.first.rfind("g1_pubkey_", 0) == 0)
      CHECK(g * Scalar::parse(kv.second[0]) == G1Point::new_(Fq1::parse(kv.second[1]), Fq1::parse(kv.second[2])));

  // ---- the MSM seam (polynomial.rs:1250-1285): 4 powers, coefficients 2..5, against the written-out sum
  {
    std::vector<G1Point> powers = {g * s(1), g * s(2), g * s(3), g * s(4)};
    Polynomial p({s(2), s(3), s(4), s(5)});
    G1Point act = p.eval_with_g1_hidings(powers);
    G1Point exp = powers[0] * s(2) + powers[1] * s(3) + powers[2] * s(4) + powers[3] * s(5);
    CHECK(act == exp);
    CHECK(act == g * s(40));
    bool threw = false;                                       // fewer powers than coefficients: the reference panics (:278)
    try { Polynomial({s(1), s(2), s(3)}).eval_with_g1_hidings({powers[0], powers[1]}); } catch (const std::runtime_error&) { threw = true; }
    CHECK(threw);
    CHECK(Polynomial({s(0)}).eval_with_g1_hidings(powers).is_zero());
    CHECK(Polynomial({s(7), s(0), s(0)}).len() == 1);         // normalize :139-152
  }

  // ---- G2 (g2_point.rs:178-289, 320-350)
  G2Point h = G2Point::g(), inf2 = G2Point::zero();
  CHECK(h * s(1) == h);
  CHECK(h * s(2) == h + h);
  CHECK(h * s(3) == h + h + h);
  {
    auto& k = kat["g2_add_same_point"];                       // x1 x0 y1 y0
    CHECK(h + h == G2Point::new_(Fq1::parse(k[0]), Fq1::parse(k[1]), Fq1::parse(k[2]), Fq1::parse(k[3])));
  }
  CHECK((h + (-h)).is_zero());
  CHECK(h + inf2 == h && inf2 + h == h && (inf2 + inf2).is_zero());
  for (int n = 1; n <= 10; n++) {
    auto& k = kat["g2_mult_" + std::to_string(n)];
    CHECK(h * s(n) == G2Point::new_(Fq1::parse(k[0]), Fq1::parse(k[1]), Fq1::parse(k[2]), Fq1::parse(k[3])));
  }
  {
    std::vector<G2Point> powers = {h * s(1), h * s(2), h * s(3)};
    CHECK(Polynomial({s(4), s(5), s(6)}).eval_with_g2_hidings(powers) == h * s(32));
  }
  std::printf("%s: %d failure(s)\n", failures ? "FAILED" : "OK", failures);
  return failures ? 1 : 0;
}
