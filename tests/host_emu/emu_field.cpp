// TEST INFRASTRUCTURE ONLY: compiles the device headers with g++ (ptx.cuh's host
// emulation of each PTX instruction) so the limb algorithms can be checked on a box
// without a GPU.  Never linked into the product library.
#include <cstring>
#include "../../zk-toolkit_b200/csrc/ec.cuh"

using namespace zk;

extern "C" {

// op: 0 add 1 sub 2 mul 3 neg 4 inv 5 to_mont 6 from_mont ; field: 0 Fq 1 Fr
int emu_field_op(int field, int op, const uint32_t* a, const uint32_t* b, uint32_t* r) {
  if (field == 0) {
    Fp x, y, z;
    memcpy(x.v, a, 48); if (b) memcpy(y.v, b, 48);
    switch (op) {
      case 0: fadd(z, x, y); break;
      case 1: fsub(z, x, y); break;
      case 2: fmul(z, x, y); break;
      case 3: fneg(z, x); break;
      case 4: finv(z, x); break;
      case 5: fto_mont(z, a); break;
      case 6: ffrom_mont(z.v, x); break;
      case 7: finv_fermat(z, x); break;
      case 8: fsqr(z, x); break;
      case 9: fmul2(z, x, y, y, y, x); fadd(z, z, y); break;   // x*y + y*x through the lockstep pair
      case 10: finv_euclid(z, x); break;
      default: return -1;
    }
    memcpy(r, z.v, 48);
  } else {
    Fr x, y, z;
    memcpy(x.v, a, 32); if (b) memcpy(y.v, b, 32);
    switch (op) {
      case 0: fadd(z, x, y); break;
      case 1: fsub(z, x, y); break;
      case 2: fmul(z, x, y); break;
      case 3: fneg(z, x); break;
      case 4: finv(z, x); break;
      case 5: fto_mont(z, a); break;
      case 6: ffrom_mont(z.v, x); break;
      case 7: finv_fermat(z, x); break;
      case 8: fsqr(z, x); break;
      case 9: fmul2(z, x, y, y, y, x); fadd(z, z, y); break;
      case 10: finv_euclid(z, x); break;
      default: return -1;
    }
    memcpy(r, z.v, 32);
  }
  return 0;
}

// Fq2 ops on Montgomery-form inputs laid out c0|c1 (24 limbs): 0 add 1 sub 2 mul 3 sqr 4 inv
int emu_fp2_op(int op, const uint32_t* a, const uint32_t* b, uint32_t* r) {
  Fp2 x, y, z;
  memcpy(&x, a, 96); if (b) memcpy(&y, b, 96);
  switch (op) {
    case 0: fadd(z, x, y); break;
    case 1: fsub(z, x, y); break;
    case 2: fmul(z, x, y); break;
    case 3: fsqr(z, x); break;
    case 4: finv(z, x); break;
    default: return -1;
  }
  memcpy(r, &z, 96);
  return 0;
}

}  // extern "C"

// G1 group ops on canonical affine inputs (24 limbs, (0,0)=inf); out canonical affine.
// op: 0 madd (P + Q via XYZZ(P) += affine Q)  1 add (XYZZ + XYZZ, Q pre-scaled)  2 dbl  3 mdbl
template <class F, int NL>
static void load_aff(Affine<F>& p, const uint32_t* in);
template <> void load_aff<Fp, 24>(Affine<Fp>& p, const uint32_t* in) {
  fto_mont(p.x, in); fto_mont(p.y, in + 12);
}
template <> void load_aff<Fp2, 48>(Affine<Fp2>& p, const uint32_t* in) {
  fto_mont(p.x.c0, in); fto_mont(p.x.c1, in + 12); fto_mont(p.y.c0, in + 24); fto_mont(p.y.c1, in + 36);
}
static void store_aff(uint32_t* out, const Affine<Fp>& p) { ffrom_mont(out, p.x); ffrom_mont(out + 12, p.y); }
static void store_aff(uint32_t* out, const Affine<Fp2>& p) {
  ffrom_mont(out, p.x.c0); ffrom_mont(out + 12, p.x.c1); ffrom_mont(out + 24, p.y.c0); ffrom_mont(out + 36, p.y.c1);
}

template <class F, int NL>
static int group_op(int op, const uint32_t* pa, const uint32_t* qa, uint32_t* out) {
  Affine<F> p, q, r;
  load_aff<F, NL>(p, pa);
  if (qa) load_aff<F, NL>(q, qa);
  XYZZ<F> acc;
  from_affine(acc, p);
  switch (op) {
    case 0: xyzz_madd(acc, q); break;
    case 1: {
      // give Q a non-trivial ZZ/ZZZ: Q = (x z^2, y z^3, z^2, z^3) with z = 5 (Montgomery)
      XYZZ<F> qq; from_affine(qq, q);
      if (!is_inf(qq)) {
        F z, z2, z3; fset_one(z); F five; fadd(five, z, z); fadd(five, five, five); fadd(five, five, z);
        fsqr(z2, five); fmul(z3, z2, five);
        fmul(qq.x, qq.x, z2); fmul(qq.y, qq.y, z3); qq.zz = z2; qq.zzz = z3;
        // and P likewise with z = 3 via two doublings' worth of scaling
        F three; fadd(three, z, z); fadd(three, three, z);
        if (!is_inf(acc)) { fsqr(z2, three); fmul(z3, z2, three); fmul(acc.x, acc.x, z2); fmul(acc.y, acc.y, z3); acc.zz = z2; acc.zzz = z3; }
      }
      xyzz_add(acc, qq);
      break;
    }
    case 2: xyzz_dbl(acc); break;
    case 3: if (!is_inf(p)) xyzz_mdbl(acc, p); break;
    default: return -1;
  }
  xyzz_to_affine(r, acc);
  store_aff(out, r);
  return 0;
}

extern "C" {
int emu_g1_op(int op, const uint32_t* p, const uint32_t* q, uint32_t* out) { return group_op<Fp, 24>(op, p, q, out); }
int emu_g2_op(int op, const uint32_t* p, const uint32_t* q, uint32_t* out) { return group_op<Fp2, 48>(op, p, q, out); }

}  // extern "C"
