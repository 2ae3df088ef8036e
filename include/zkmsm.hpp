// zkmsm.hpp -- C++ host-side mirror of the reference's interface on the MSM path, over the C ABI of zkmsm.h.
//
// The reference is compiled code (Rust); its toolchain is not in this image, so the host layer a Rust
// maintainer would write is mirrored here in C++ with the same names, argument meaning and
// error behaviour:
//
//   reference (src/...)                                                     here (namespace zk_toolkit)
//   building_block/curves/bls12_381/g1_point.rs:33-36   enum G1Point        G1Point  (Rational{x,y} | AtInfinity)
//   building_block/curves/bls12_381/g2_point.rs:31-34   enum G2Point        G2Point
//   building_block/curves/macros.rs:35-163              impl Add            operator+
//   building_block/curves/macros.rs:2-32                impl Mul<&Fq1>      operator*(point, Scalar)
//   g1_point.rs:178-195                                 impl Neg            operator-
//   g1_point.rs:163-174                                 impl PartialEq      operator==
//   building_block/field/polynomial.rs:272-293          eval_with_g{1,2}_hidings   Polynomial::eval_with_g{1,2}_hidings
//
// Failures throw std::runtime_error where the reference panics (polynomial.rs:278 index out of bounds;
// unwrap()s), never return a wrong value.  There is no CPU arithmetic here: every group operation is a call into
// libzkmsm.so, and constructing the Gpu context fails without an sm_100 device.
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "zkmsm.h"

namespace zk_toolkit {

// fixed-width little-endian integers at the ABI: N 32-bit limbs
template <size_t N> struct Limbs {
  std::array<uint32_t, N> w{};
  bool operator==(const Limbs& o) const { return w == o.w; }
  bool operator!=(const Limbs& o) const { return !(w == o.w); }
  bool is_zero() const { for (auto v : w) if (v) return false; return true; }
  static Limbs from_u64(uint64_t v) { Limbs r; r.w[0] = (uint32_t)v; if (N > 1) r.w[1] = (uint32_t)(v >> 32); return r; }
  // decimal or 0x-hex literal, as the reference's tests write their constants (BigUint::parse_bytes)
  static Limbs parse(const std::string& s) {
    Limbs r;
    size_t i = 0;
    uint32_t base = 10;
    if (s.size() > 2 && s[0] == '0' && (s[1] == 'x' || s[1] == 'X')) { base = 16; i = 2; }
    for (; i < s.size(); i++) {
      char c = s[i];
      uint32_t d = c >= '0' && c <= '9' ? c - '0' : c >= 'a' && c <= 'f' ? c - 'a' + 10 : c >= 'A' && c <= 'F' ? c - 'A' + 10 : 99;
      if (d >= base) throw std::runtime_error("bad digit in integer literal");
      uint64_t carry = d;
      for (size_t k = 0; k < N; k++) { uint64_t t = (uint64_t)r.w[k] * base + carry; r.w[k] = (uint32_t)t; carry = t >> 32; }
      if (carry) throw std::runtime_error("integer literal too large");
    }
    return r;
  }
};
typedef Limbs<12> Fq1;     // base-field element, canonical (< q)
typedef Limbs<8> Scalar;   // Fr element / raw multiplier

class Gpu {  // one device context, created on first use (MclInitializer::init precedent, mcl_initializer.rs:8-15)
 public:
  static zkmsm_ctx* ctx() {
    static Gpu g;
    return g.ctx_;
  }
  static void check(int rc) {
    if (rc != ZKMSM_OK) throw std::runtime_error(std::string("zkmsm: ") + zkmsm_last_error(ctx()));
  }
 private:
  Gpu() {
    int rc = zkmsm_create(0, &ctx_);
    if (rc != ZKMSM_OK) throw std::runtime_error("zkmsm_create failed (an sm_100 CUDA device is required; there is no CPU path)");
  }
  ~Gpu() { zkmsm_destroy(ctx_); }
  zkmsm_ctx* ctx_ = nullptr;
};

template <size_t WORDS, int GROUP> struct Point {
  bool at_infinity = true;
  std::array<uint32_t, WORDS> xy{};   // G1: x | y ; G2: x.u0 | x.u1 | y.u0 | y.u1 (canonical limbs)

  static Point zero() { return Point(); }   // Zero::zero(), g1_point.rs:131-142
  bool is_zero() const { return at_infinity; }
  bool operator==(const Point& o) const { return at_infinity == o.at_infinity && (at_infinity || xy == o.xy); }
  bool operator!=(const Point& o) const { return !(*this == o); }

  static Point from_abi(const uint32_t* w, int inf) {
    Point p;
    p.at_infinity = inf != 0;
    if (!inf) std::memcpy(p.xy.data(), w, WORDS * 4);
    return p;
  }

  // &self + &rhs (impl_affine_add!): the 2-term MSM 1*self + 1*rhs
  Point operator+(const Point& rhs) const {
    uint32_t pts[2 * WORDS], sc[16] = {1, 0, 0, 0, 0, 0, 0, 0, 1, 0, 0, 0, 0, 0, 0, 0}, out[WORDS];
    uint8_t inf[2] = {(uint8_t)at_infinity, (uint8_t)rhs.at_infinity};
    std::memcpy(pts, xy.data(), WORDS * 4);
    std::memcpy(pts + WORDS, rhs.xy.data(), WORDS * 4);
    int oinf = 0;
    Gpu::check(GROUP == 1 ? zkmsm_g1_msm_oneshot(Gpu::ctx(), pts, inf, sc, 2, out, &oinf)
                          : zkmsm_g2_msm_oneshot(Gpu::ctx(), pts, inf, sc, 2, out, &oinf));
    return from_abi(out, oinf);
  }
  // &self * &k (impl_scalar_mul_point!): raw integer multiple, not reduced mod r
  Point operator*(const Scalar& k) const {
    if (at_infinity) return Point();
    uint32_t out[WORDS];
    uint8_t oinf = 0;
    Gpu::check(GROUP == 1 ? zkmsm_g1_mul_base(Gpu::ctx(), xy.data(), k.w.data(), 1, out, &oinf)
                          : zkmsm_g2_mul_base(Gpu::ctx(), xy.data(), k.w.data(), 1, out, &oinf));
    return from_abi(out, oinf);
  }
};

struct G1Point : Point<24, 1> {
  G1Point() {}
  G1Point(const Point<24, 1>& p) : Point<24, 1>(p) {}
  static G1Point new_(const Fq1& x, const Fq1& y) {   // G1Point::new, g1_point.rs:50-55
    G1Point p;
    p.at_infinity = false;
    std::memcpy(p.xy.data(), x.w.data(), 48);
    std::memcpy(p.xy.data() + 12, y.w.data(), 48);
    return p;
  }
  static G1Point g() {   // g1_point.rs:38-47
    return new_(Fq1::parse("0x17f1d3a73197d7942695638c4fa9ac0fc3688c4f9774b905a14e3a3f171bac586c55e83ff97a1aeffb3af00adb22c6bb"),
                Fq1::parse("0x08b3f481e3aaa0f1a09e30ed741d8ae4fcf5e095d5d00af600db18cb2c04b3edd03cc744a2888ae40caa232946c5e7e1"));
  }
  Fq1 x() const { Fq1 r; std::memcpy(r.w.data(), xy.data(), 48); return r; }
  Fq1 y() const { Fq1 r; std::memcpy(r.w.data(), xy.data() + 12, 48); return r; }
  G1Point operator-() const;   // Neg, g1_point.rs:178-195
};

struct G2Point : Point<48, 2> {
  G2Point() {}
  G2Point(const Point<48, 2>& p) : Point<48, 2>(p) {}
  // Fq2::new takes (u1, u0) (fq2.rs:22); the ABI stores u0 first
  static G2Point new_(const Fq1& x_u1, const Fq1& x_u0, const Fq1& y_u1, const Fq1& y_u0) {
    G2Point p;
    p.at_infinity = false;
    const Fq1* order[4] = {&x_u0, &x_u1, &y_u0, &y_u1};
    for (int k = 0; k < 4; k++) std::memcpy(p.xy.data() + 12 * k, order[k]->w.data(), 48);
    return p;
  }
  static G2Point g() {   // g2_point.rs:36-46
    return new_(Fq1::parse("0x13e02b6052719f607dacd3a088274f65596bd0d09920b61ab5da61bbdc7f5049334cf11213945d57e5ac7d055d042b7e"),
                Fq1::parse("0x024aa2b2f08f0a91260805272dc51051c6e47ad4fa403b02b4510b647ae3d1770bac0326a805bbefd48056c8c121bdb8"),
                Fq1::parse("0x0606c4a02ea734cc32acd2b02bc28b99cb3e287e85a763af267492ab572e99ab3f370d275cec1da1aaa9075ff05f79be"),
                Fq1::parse("0x0ce5d527727d6e118cc9cdc6da2e351aadfd9baa8cbdd3a76d429a695160d12c923ac9cc3baca289e193548608b82801"));
  }
  G2Point operator-() const;
};

namespace detail {
// q - v for a canonical v (and 0 -> 0: prime_field_elem.rs:448-457); the one piece of host arithmetic, a 12-limb subtraction
inline void negate_mod_q(uint32_t* v) {
  static const Fq1 q = Fq1::parse("0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab");
  bool zero = true;
  for (int i = 0; i < 12; i++) zero = zero && v[i] == 0;
  if (zero) return;
  uint64_t borrow = 0;
  for (int i = 0; i < 12; i++) {
    uint64_t t = (uint64_t)q.w[i] - v[i] - borrow;
    v[i] = (uint32_t)t;
    borrow = (t >> 32) & 1;
  }
}
}  // namespace detail

inline G1Point G1Point::operator-() const {
  G1Point r = *this;
  if (!at_infinity) detail::negate_mod_q(r.xy.data() + 12);
  return r;
}
inline G2Point G2Point::operator-() const {
  G2Point r = *this;
  if (!at_infinity) { detail::negate_mod_q(r.xy.data() + 24); detail::negate_mod_q(r.xy.data() + 36); }
  return r;
}

// field/polynomial.rs:32-36 restricted to the MSM seam: coefficients in Fr, coeffs[i] multiplies x^i
class Polynomial {
 public:
  explicit Polynomial(std::vector<Scalar> coeffs) : coeffs_(std::move(coeffs)) {
    if (coeffs_.empty()) throw std::runtime_error("coeffs is empty");                     // polynomial.rs:120
    while (coeffs_.size() > 1 && coeffs_.back().is_zero()) coeffs_.pop_back();             // normalize, :139-152
  }
  size_t len() const { return coeffs_.size(); }

  G1Point eval_with_g1_hidings(const std::vector<G1Point>& powers) const {   // polynomial.rs:272-281
    return G1Point(eval<24, 1>(powers));
  }
  G2Point eval_with_g2_hidings(const std::vector<G2Point>& powers) const {   // polynomial.rs:284-293
    return G2Point(eval<48, 2>(powers));
  }

 private:
  template <size_t WORDS, int GROUP, class P> Point<WORDS, GROUP> eval(const std::vector<P>& powers) const {
    const size_t n = coeffs_.size();
    if (powers.size() < n) throw std::runtime_error("index out of bounds: more coefficients than powers");   // :278
    std::vector<uint32_t> xy(n * WORDS), sc(n * 8);
    std::vector<uint8_t> inf(n);
    for (size_t i = 0; i < n; i++) {
      std::memcpy(&xy[i * WORDS], powers[i].xy.data(), WORDS * 4);
      inf[i] = powers[i].at_infinity;
      std::memcpy(&sc[i * 8], coeffs_[i].w.data(), 32);
    }
    uint32_t out[WORDS];
    int oinf = 0;
    Gpu::check(GROUP == 1 ? zkmsm_g1_msm_oneshot(Gpu::ctx(), xy.data(), inf.data(), sc.data(), n, out, &oinf)
                          : zkmsm_g2_msm_oneshot(Gpu::ctx(), xy.data(), inf.data(), sc.data(), n, out, &oinf));
    return Point<WORDS, GROUP>::from_abi(out, oinf);
  }
  std::vector<Scalar> coeffs_;
};

}  // namespace zk_toolkit
